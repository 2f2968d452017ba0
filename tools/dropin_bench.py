#!/usr/bin/env python
"""Cost of the 9-argument drop-in symbol message_passing() of libldpc_b200.so (one frame per call, as the reference's
frame loops call it, parallel_simulator.py:152-164): tight ctypes loop with preallocated arrays.
  fixed code  : the same lookups every call (handle found by sampled hash + memcmp)
  ensemble    : fresh lookups every call (graph upload + validation per call; the reference spends 15 ms per graph here)
usage: python tools/dropin_bench.py [n] [seconds]"""
import ctypes as ct
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iib_project_ldpc_codes_b200 as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
dv, dc, its, eps = 3, 6, 200, 0.40
k = n - n * dv // dc
lib = ct.CDLL(L.LIB_PATH)
P = ct.POINTER(ct.c_int)
rng = np.random.default_rng(1)


def loop(graphs, seconds):
    errors = np.zeros(its, dtype=np.int32)
    frames, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        vl, cl = graphs[frames % len(graphs)]
        mvc = np.where(rng.random(n) < eps, 2, 0).astype(np.int32)
        errors[:] = 0
        r = lib.message_passing(mvc.ctypes.data_as(P), its, vl.ctypes.data_as(P), cl.ctypes.data_as(P), errors.ctypes.data_as(P), n, k, dv, dc)
        assert r >= 0, L._lib.load().ldpc_b200_last_error()
        frames += 1
    return frames / (time.perf_counter() - t0)


g1 = [tuple(np.ascontiguousarray(a.reshape(-1), dtype=np.int32) for a in L.generate_graph(n, dv, dc, seed=1, mode=1))]
loop(g1, 0.5)
print(f"n={n} fixed code : {loop(g1, secs):9.0f} frames/s through message_passing() (one GPU launch per frame)")
many = [tuple(np.ascontiguousarray(a.reshape(-1), dtype=np.int32) for a in L.generate_graph(n, dv, dc, seed=100 + s, mode=1)) for s in range(64)]
loop(many, 0.5)
print(f"n={n} ensemble   : {loop(many, secs):9.0f} frames/s (64 graphs in rotation: 16-entry handle cache misses every call)")

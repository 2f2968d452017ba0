#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference C compiled in the build container.

Run where /root/reference exists:  python tests/golden/make_golden.py
It builds oracle/_ref (gcc on /root/reference/{message_passing,random_code_generator}.c where they lie)
and records inputs + the reference's outputs, so that machines without /root/reference (the GPU box)
can still pin the oracle restatement and the CUDA path against the reference's behaviour.

Files
  mp_cases.npz   single-frame calls of message_passing.c:7 through oracle/_ref/message_passing.so:
                 graphs from random_code_generator.c:21 (libc srand(1)), random 0/1/2 inputs incl. invalid
                 codewords, iteration caps {0,1,2,3,8,50}; hand-built multi-edge graph, all 81 inputs
  mc_cases.npz   counter blocks of the reference frame loop (oracle/ref_harness.c around the reference
                 object) on the erasure patterns of oracle/bec_rng.h, incl. expurgation
  rng_words.npz  raw words of the host restatement of the erasure generator (oracle/bec_rng.h) -- pins the
                 RNG definition itself against accidental edits (not a reference artefact: the reference
                 draws from numpy's unseeded global generator, channels.py:24-26)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    O.build()
    assert O.have_ref(), "needs /root/reference to build oracle/_ref"
    rng = np.random.default_rng(20261018)
    out = {}
    idx = 0

    def add_case(v2c, c2v, n, k, dv, dc, mvc, its, prefill=None):
        nonlocal idx
        M, err, ret = O.single_frame("ref", mvc, its, v2c, c2v, n, k, dv, dc, errors=prefill)
        p = f"c{idx:03d}_"
        out[p + "dims"] = np.array([n, k, dv, dc, its], dtype=np.int32)
        out[p + "v2c"] = np.asarray(v2c, dtype=np.int32).ravel()
        out[p + "c2v"] = np.asarray(c2v, dtype=np.int32).ravel()
        out[p + "in"] = np.asarray(mvc, dtype=np.int8)
        out[p + "prefill"] = np.zeros(its, np.int32) if prefill is None else np.asarray(prefill, np.int32)
        out[p + "out"] = M.astype(np.int8)
        out[p + "errors"] = err.astype(np.int32)
        out[p + "ret"] = np.array([ret], dtype=np.int32)
        idx += 1

    # the commented-out main() of message_passing.c:180-189
    v2c = [0, 1, 0, 1, 0, 2, 1, 2, 0, 2, 1, 2]
    c2v = [0, 1, 2, 4, 0, 1, 3, 5, 2, 3, 4, 5]
    for mvc, its in [([1, 1, 2, 0, 1, 0], 1), ([1, 1, 2, 0, 1, 0], 5), ([2, 2, 2, 2, 1, 0], 6), ([2, 2, 0, 0, 1, 0], 1),
                     ([0] * 6, 3), ([2, 0, 1, 0, 0, 0], 3), ([2, 2, 2, 2, 2, 2], 4), ([1, 1, 2, 0, 1, 0], 0)]:
        add_case(v2c, c2v, 6, 3, 2, 4, mvc, its)
    add_case(v2c, c2v, 6, 3, 2, 4, [2, 2, 2, 2, 1, 0], 6, prefill=[100] * 6)
    add_case(v2c, c2v, 6, 3, 2, 4, [2, 2, 2, 2, 1, 0], 6, prefill=[3, 1, 0, 0, 2, 2])
    v2c_sw = list(v2c); v2c_sw[0:2] = [1, 0]
    add_case(v2c_sw, c2v, 6, 3, 2, 4, [2, 0, 1, 0, 0, 0], 3)
    # hand-built multi-edge graph, every input in {0,1,2}^4
    mc2v = [0, 0, 1, 2, 1, 2, 3, 3]
    mv2c = [0, 0, 0, 1, 0, 1, 1, 1]
    for a in range(81):
        add_case(mv2c, mc2v, 4, 2, 2, 4, [a // 27 % 3, a // 9 % 3, a // 3 % 3, a % 3], 4)
    # graphs from the reference generator
    for (n, dv, dc, eps) in [(24, 3, 6, 0.4), (96, 3, 6, 0.42), (200, 4, 8, 0.38), (60, 3, 4, 0.5), (1000, 3, 6, 0.42)]:
        vl, cl, _ = O.ref_generate_random_code(n, dv, dc, srand_seed=1)
        k = n - n * dv // dc
        reps = 4 if n >= 1000 else 12
        for r in range(reps):
            vals = rng.integers(0, 2, n) if r % 2 else np.zeros(n, dtype=np.int64)
            mvc = np.where(rng.random(n) < eps, 2, vals)
            add_case(vl, cl, n, k, dv, dc, mvc, [0, 1, 2, 3, 8, 50][r % 6])
    out["num_cases"] = np.array([idx], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "mp_cases.npz"), **out)
    print("mp_cases:", idx, "cases")

    # Monte-Carlo counter blocks through the reference object
    mc = {}
    j = 0
    for (n, dv, dc, eps, frames, its, exp, first) in [
            (96, 3, 6, [0.30, 0.45], 500, 30, -1, 0), (96, 3, 6, [0.40], 333, 12, 0, 128), (96, 3, 6, [0.40], 333, 12, 7, 0),
            (1000, 3, 6, [0.40, 0.43], 300, 200, -1, 0), (1000, 3, 6, [0.41], 300, 200, 0, 256),
            (200, 4, 8, [0.36], 400, 50, -1, 0), (96, 3, 6, [0.0, 1.0], 64, 5, -1, 0), (96, 3, 6, [0.4], 100, 0, -1, 0),
            (96, 3, 6, [0.4], 100, 1, -1, 0), (96, 3, 6, [0.4], 100, 2, -1, 0)]:
        vl, cl, _ = O.ref_generate_random_code(n, dv, dc, srand_seed=1)
        k = n - n * dv // dc
        seed = 1000 + j
        c = O.reference().mc_run(vl, cl, n, k, dv, dc, eps, seed, first, frames, its, exp, nthreads=4)
        p = f"m{j:02d}_"
        mc[p + "dims"] = np.array([n, k, dv, dc, its, exp, frames, first, seed], dtype=np.int64)
        mc[p + "eps"] = np.array(eps, dtype=np.float64)
        mc[p + "v2c"] = vl.ravel().astype(np.int32)
        mc[p + "c2v"] = cl.ravel().astype(np.int32)
        mc[p + "counters"] = c
        j += 1
    mc["num_cases"] = np.array([j], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "mc_cases.npz"), **mc)
    print("mc_cases:", j, "cases")

    words = {}
    pts = [(0x0123456789ABCDEF, 0, 0, 0.40), (1, 3, 2 ** 33 + 5, 0.43), (12345, 1, 7, 0.5), (99, 0, 1, 1e-6),
           (7, 2, 9, 0.999), (5, 0, 0, 0.3834), (5, 0, 0, 0.25), (5, 0, 0, 2.0 ** -32)]
    for t, (seed, point, q, eps) in enumerate(pts):
        words[f"w{t}_key"] = np.array([seed, point, q], dtype=np.uint64)
        words[f"w{t}_eps"] = np.array([eps])
        words[f"w{t}_words"] = np.array([O.bec_word(seed, point, q, v, eps) for v in range(64)], dtype=np.uint32)
    words["num"] = np.array([len(pts)], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "rng_words.npz"), **words)
    print("rng_words:", len(pts))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Randomised differential campaign: CUDA path (all engines / geometries) vs the oracle on random codes.

usage: python tools/fuzz.py [seconds] [seed]
Every case draws a random (n, dv, dc), iteration cap, erasure points, frame count, expurgation threshold,
engine and word width, and compares the fused Monte-Carlo counters and a batched decode bit for bit."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iib_project_ldpc_codes_b200 as L  # noqa: E402
from oracle import oracle as O  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
chk = O.best()
t0 = time.time()
cases = fails = 0
DEGS = [(2, 4), (3, 6), (4, 8), (3, 4), (3, 9), (5, 10), (2, 6), (3, 12), (4, 6)]
while time.time() - t0 < budget:
    dv, dc = DEGS[rng.integers(len(DEGS))]
    step = dc // np.gcd(dv, dc)
    n = int(rng.choice([step * rng.integers(2, 40), step * rng.integers(40, 400), step * rng.integers(400, 1500)]))
    n = max(n, dc + step)
    n -= n % step
    try:
        vl, cl = L.generate_graph(n, dv, dc, seed=int(rng.integers(1 << 30)), mode=1)
    except RuntimeError:
        continue
    m = n * dv // dc
    its = int(rng.choice([0, 1, 2, 3, 5, 17, 60, 200]))
    npts = int(rng.integers(1, 4))
    thr = 1 - (1 - 1.0 / (dc - 1)) if dv == 2 else 0.45       # rough scale of interesting eps
    eps = [float(np.clip(rng.normal(0.8 * thr, 0.15), 0.0, 1.0)) for _ in range(npts)]
    frames = int(rng.choice([1, 31, 32, 33, 127, 128, 129, 500, 1000, 2500]))
    if rng.random() < (0.4 if (dv, dc) in ((3, 6), (4, 8)) else 0.15):
        frames = int(rng.integers(33000, 36000))            # long enough for the staged path (and the overlapped stage 0)
    if n * frames * max(its, 1) > 6e9:
        frames = max(1, int(6e9 / (n * max(its, 1))))
    exp = int(rng.choice([-1, -1, 0, 1, n // 20, n]))
    engine = int(rng.choice([0, 0, 1, 2]))
    words = int(rng.choice([0, 1, 2, 4]))
    first = int(rng.integers(0, 50)) * 128
    seed = int(rng.integers(1 << 62))
    # round-2 paths, switched per case: overlapped stage 0 (auto / forced / off) with 1..8 producer warps, socket lists in
    # tensor or shared memory, E2 lookups through the TMA or plain loads, two-pass expurgation forced, blocking or
    # background placement search
    knobs = {"LDPC_B200_OVERLAP": int(rng.choice([0, 1, 1, 2])), "LDPC_B200_PRODUCERS": int(rng.choice([1, 2, 3, 4, 8])),
             "LDPC_B200_TMEM": int(rng.choice([0, 0, 2])), "LDPC_B200_E2_TMA": int(rng.choice([1, 2])),
             "LDPC_B200_EXP2PASS": int(rng.choice([0, 1])), "LDPC_B200_PLACE_SYNC": int(rng.choice([0, 1])),
             "LDPC_B200_CACHE": 0}
    for k_, v_ in knobs.items():
        os.environ[k_] = str(v_)
    desc = (f"n={n} ({dv},{dc}) its={its} eps={np.round(eps, 3).tolist()} frames={frames} exp={exp} engine={engine} words={words} "
            f"first={first} knobs={knobs}")
    try:
        g = L.Graph(vl, cl)
        g.set_engine(engine)
        g.set_tuning(words, 0, int(rng.choice([0, 0, 1, 2, 3])))
        g.set_place_effort(int(rng.choice([0, 1, 2])))
        try:
            got, st = g.mc_run(eps, seed, frames, its, expurgation=exp, first_frame=first, return_stats=True)
        except L.LdpcB200Error as e:
            if "does not fit" in str(e) and engine == 1:
                continue
            raise
        want = chk.mc_run(vl, cl, n, n - m, dv, dc, eps, seed, first, frames, its, exp, nthreads=16)
        ok = np.array_equal(got, want)
        # batched decode with values on a few frames
        fr = int(rng.integers(1, 70))
        M = np.where(rng.random((fr, n)) < eps[0], 2, rng.integers(0, 2, (fr, n))).astype(np.int8)
        try:
            a = g.decode_batch(M, its)
            b = chk.decode_batch(vl, cl, n, n - m, dv, dc, M, its)
            ok = ok and all(np.array_equal(x, y) for x, y in zip(a, b))
        except L.LdpcB200Error as e:
            if not ("does not fit" in str(e) and engine == 1):
                raise
        g.close()
    except Exception as e:  # noqa: BLE001
        ok = False
        desc += f" EXC {type(e).__name__}: {e}"
    cases += 1
    if not ok:
        fails += 1
        print("MISMATCH", desc, flush=True)
print(f"fuzz: {cases} cases, {fails} failures, {time.time() - t0:.0f} s", flush=True)
sys.exit(1 if fails else 0)

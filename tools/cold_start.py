#!/usr/bin/env python
"""Time line of a cold start at BASELINE config 2 (n = 10^4, six eps points, 10^7 frames per point): first, second and third
host-buffer call on a fresh handle with the placement cache disabled; LDPC_B200_TRACE prints which layout (natural /
quick placement / full placement) every frame chunk was enqueued on and when."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LDPC_B200_CACHE"] = "0"; os.environ["LDPC_B200_TRACE"] = "1"
t00 = time.time()
import numpy as np, torch
import iib_project_ldpc_codes_b200 as L
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
vl, cl = L.generate_graph(10000, 3, 6, seed=1, mode=1)
eps = [0.38, 0.39, 0.40, 0.41, 0.42, 0.43]
t0 = time.time(); g = L.Graph(vl, cl); t1 = time.time()
c = g.mc_run(eps, 1, 10_000_000, 200); t2 = time.time()
print(f"graph create {t1-t0:.3f} s, first mc_run {t2-t1:.3f} s", flush=True)
c = g.mc_run(eps, 2, 10_000_000, 200); t3 = time.time()
print(f"second mc_run {t3-t2:.3f} s", flush=True)
g.wait_placement(); t4 = time.time()
c = g.mc_run(eps, 3, 10_000_000, 200); t5 = time.time()
print(f"wait {t4-t3:.3f} s, third mc_run {t5-t4:.3f} s", flush=True)

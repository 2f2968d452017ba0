#!/usr/bin/env python
"""BASELINE config 0 as the reference runs it: (3,6) n=1000, eps=0.40, 10^4 frames, cap 200 -- wall time of one host call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import iib_project_ldpc_codes_b200 as L
vl, cl = L.generate_graph(1000, 3, 6, seed=1, mode=1)
t = time.time(); g = L.Graph(vl, cl); c = g.mc_run([0.40], 1, 10000, 200); t_first = time.time() - t
ts = []
for s in range(20):
    t = time.time(); c = g.mc_run([0.40], 2 + s, 10000, 200); ts.append(time.time() - t)
fer = c[0, 201] / c[0, 203]
print(f"first call (context, placement, kernel load): {t_first:.3f} s; steady state: median {np.median(ts)*1e3:.3f} ms per 10^4-frame call "
      f"= {1e4/np.median(ts):.3g} frames/s; FER {fer:.4f}")

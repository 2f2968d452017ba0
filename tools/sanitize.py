#!/usr/bin/env python
"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize.py
Each case is also checked against the oracle when it is importable."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iib_project_ldpc_codes_b200 as L  # noqa: E402

vl, cl = L.generate_graph(504, 3, 6, seed=2, mode=1)
g = L.Graph(vl, cl)
a = g.mc_run([0.40, 0.44], 5, 3000, 40)                      # unstaged E1 (W=128 teams), draw queue
b = g.mc_run([0.40], 5, 3000, 40, expurgation=1)             # masked second pass
os.environ["LDPC_B200_STAGES"] = "3"
c = L.Graph(vl, cl).mc_run([0.40, 0.44], 5, 40000, 40)       # staged E1: spill / transposes / re-pack
os.environ.pop("LDPC_B200_STAGES")
for w in (1, 2):
    h = L.Graph(vl, cl); h.set_tuning(w, 0, 0)
    h.mc_run([0.42], 6, 2000, 30)
pats = L.draw_patterns(504, 0.42, 7, 0, 0, 70)
d = g.decode_batch(pats, 30)                                  # batched decode (value plane)
e = L.mc_run_ensemble(504, 3, 6, [0.40], 3, 40, 40, frames_per_graph=128, gen_mode=2, expurgation=0)   # in-kernel graphs
v2, c2 = L.generate_graph(20000, 3, 6, seed=1, mode=1)
g2 = L.Graph(v2, c2)
f = g2.mc_run([0.40], 9, 2048, 30)                            # E2 (HBM/L2 streamed, cooperative)
v3, c3 = L.generate_graph(10000, 3, 6, seed=1, mode=1)
g3 = L.Graph(v3, c3)
k = g3.mc_run([0.41], 9, 4096, 30)                            # E1 W=64, L2-streamed + resident variable-side lists
print("sanitize cases done", int(a.sum()), int(b.sum()), int(c.sum()), int(e.sum()), int(f.sum()), int(k.sum()))

#!/usr/bin/env python
"""Finite-length sweep of BASELINE config 5: (3,6)-regular n in {500, 1000, 2000, 4000, 8000}, 20 eps points 0.34-0.45,
frames per point scaled so that every n does the same number of edge updates; plus (4,8) n=2000 without expurgation.
Prints frames/s, edge-updates/s and the fraction of the measured shared-memory roofline per n."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import iib_project_ldpc_codes_b200 as L

peak = L.api.smem_bandwidth_gbs(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
eps = [round(0.34 + 0.11 * i / 19, 6) for i in range(20)]
cases = [(n, 3, 6, eps, 4_000_000_000 // (20 * n)) for n in (500, 1000, 2000, 4000, 8000)]
cases.append((2000, 4, 8, [0.30, 0.32, 0.34, 0.36], 25_000_000))
for n, dv, dc, e, frames in cases:
    vl, cl = L.generate_graph(n, dv, dc, seed=1, mode=1)
    g = L.Graph(vl, cl)
    d = torch.zeros((len(e), 206), dtype=torch.int64, device=dev)
    g.mc_launch(e, 1, frames, 200, d.data_ptr(), st); torch.cuda.synchronize()
    g.wait_placement()   # the first launch overlaps the background placement search (natural layout)
    g.mc_launch(e, 3, frames, 200, d.data_ptr(), st); torch.cuda.synchronize()
    g.kernel_times_ms(); g.mc_stats(); d.zero_()
    g.mc_launch(e, 2, frames, 200, d.data_ptr(), st); torch.cuda.synchronize()
    ms = float(g.kernel_times_ms()[0]); s = g.mc_stats()
    c = d.cpu().numpy().view(np.uint64)
    ex, fr = int(c[:, 205].sum()), int(c[:, 203].sum())
    b = (2 + 1 / dc + 2 / dv) / 8 + 2 * 2 / 128
    eu = ex * n * dv / ms * 1e3
    print(f"n={n} ({dv},{dc}) points={len(e)} frames/point={frames}: {fr/ms*1e3:.4g} frames/s  {eu:.4g} EU/s  "
          f"{eu*b/1e9/peak:.3f} of smem roofline  W={32*s['words']} teams={s['teams']}x{s['team_warps']} stages={s['stages']} "
          f"inflation={s['phys_frame_iterations']/ex:.3f}", flush=True)

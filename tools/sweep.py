#!/usr/bin/env python
"""Tuning sweep of the fused Monte-Carlo kernel: frames/s for (words, teams, team_warps) choices.
usage: python tools/sweep.py n dv dc eps frames [words,teams,team_warps ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import iib_project_ldpc_codes_b200 as L  # noqa: E402

n, dv, dc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
eps = [float(x) for x in sys.argv[4].split(",")]
frames = int(sys.argv[5])
cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[6:]] or [(0, 0, 0)]
cap = 200
vl, cl = L.generate_graph(n, dv, dc, seed=1, mode=1)
g = L.Graph(vl, cl)
dev = torch.device("cuda", 0)
d = torch.zeros((len(eps), cap + 6), dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for cfg in cfgs:
    g.set_tuning(*cfg)
    try:
        g.mc_launch(eps, 1, frames, cap, d.data_ptr(), st)
        torch.cuda.synchronize()
        g.wait_placement()   # the first launch may have run on the natural layout (background placement search)
        g.mc_launch(eps, 3, min(frames, 200000), cap, d.data_ptr(), st)
        torch.cuda.synchronize()
        g.kernel_times_ms(); g.mc_stats()
        d.zero_()
        g.mc_launch(eps, 2, frames, cap, d.data_ptr(), st)
        torch.cuda.synchronize()
        ms = float(g.kernel_times_ms()[0])
        s = g.mc_stats()
        c = d.cpu().numpy().view(np.uint64)
        ex = int(c[:, cap + 5].sum()); fr = int(c[:, cap + 3].sum())
        print(f"cfg={cfg} words={s['words']} teams={s['teams']} tw={s['team_warps']} grid={s['grid']} smem={s['smem_bytes']} "
              f"ms={ms:.2f} frames/s={fr/ms*1e3:.4g} EU/s={ex*n*dv/ms*1e3:.4g} it/frame={ex/fr:.2f} "
              f"inflation={s['phys_frame_iterations']/ex:.3f} engine={s['engine']} stages={s['stages']} launches={s['kernel_launches']} "
              f"place={g.placement_info(s['words']) if s['engine'] == 1 else None}"
              + (f" overlap: producers={s['overlap_producers']} Mcycles/CTA cons_wait={s['overlap_cycles'][0]/s['grid']/1e6:.2f} "
                 f"prod_wait={s['overlap_cycles'][1]/s['grid']/1e6:.2f} draw={s['overlap_cycles'][2]/s['grid']/1e6:.2f} "
                 f"decode={s['overlap_cycles'][3]/s['grid']/1e6:.2f}" if s.get('overlap_producers') else ""), flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"cfg={cfg} failed: {e}", flush=True)

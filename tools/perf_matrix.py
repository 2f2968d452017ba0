#!/usr/bin/env python
"""Fixed performance matrix of the fused Monte-Carlo path (one GPU): the same workloads and frame counts every time,
so that two builds can be compared line by line.  usage: python tools/perf_matrix.py [TAG] [names...]
Prints one line per workload: frames/s, edge-updates/s, fraction of the shared-memory roofline (37.0 TB/s nominal at
1965 MHz unless LDPC_PEAK_TBS is set), lock-step inflation, geometry."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import iib_project_ldpc_codes_b200 as L  # noqa: E402

MATRIX = {
    # name: (n, dv, dc, eps, frames per point, expurgation)
    "c2": (10000, 3, 6, [0.38, 0.39, 0.40, 0.41, 0.42, 0.43], 2_000_000, -1),
    "c1": (1000, 3, 6, [0.40], 40_000_000, -1),
    "n500": (500, 3, 6, [0.36, 0.40, 0.42], 8_000_000, -1),
    "n2000": (2000, 3, 6, [0.36, 0.40, 0.42], 4_000_000, -1),
    "n4000": (4000, 3, 6, [0.36, 0.40, 0.42], 2_000_000, -1),
    "n8000": (8000, 3, 6, [0.36, 0.40, 0.42], 1_000_000, -1),
    "c4fix": (2000, 4, 8, [0.30, 0.32, 0.34, 0.36], 5_000_000, -1),
    "c4exp": (2000, 4, 8, [0.30, 0.32, 0.34, 0.36], 5_000_000, 0),
}
PEAK = float(os.environ.get("LDPC_PEAK_TBS", "37.0")) * 1e12
tag = sys.argv[1] if len(sys.argv) > 1 else "perf"
names = sys.argv[2:] or list(MATRIX)
cap = 200
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
for name in names:
    n, dv, dc, eps, frames, exp = MATRIX[name]
    vl, cl = L.generate_graph(n, dv, dc, seed=1, mode=1)
    g = L.Graph(vl, cl)
    d = torch.zeros((len(eps), cap + 6), dtype=torch.int64, device=dev)
    best = None
    for rep in range(3):   # first = warm-up (placement, pilot), then best of two
        d.zero_()
        g.kernel_times_ms(); g.mc_stats()
        g.mc_launch(eps, 100 + rep, frames, cap, d.data_ptr(), st, expurgation=exp)
        torch.cuda.synchronize()
        ms = float(g.kernel_times_ms()[0])
        s = g.mc_stats()
        if rep and (best is None or ms < best[0]):
            c = d.cpu().numpy().view(np.uint64)
            best = (ms, s, int(c[:, cap + 5].sum()), int(c[:, cap + 3].sum()))
    ms, s, ex, fr = best
    balg = (2 + 1 / dc + 2 / dv) / 8 + 2 * 2 / 128
    eu = ex * n * dv / ms * 1e3
    print(f"[{tag}] {name:6s} frames/s={fr / ms * 1e3:.4g} EU/s={eu:.4g} frac={eu * balg / PEAK:.3f} ms={ms:.2f} "
          f"infl={s['phys_frame_iterations'] / max(ex, 1):.3f} W={32 * s['words']} teams={s['teams']}x{s['team_warps']} "
          f"smem={s['smem_bytes']} stages={s['stages']} launches={s['kernel_launches']} 2nd={s['second_pass_groups']}", flush=True)
    g.close()

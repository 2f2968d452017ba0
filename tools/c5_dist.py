#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: finite-length scaling sweep n in {500, 1000, 2000, 4000, 8000} x 20 eps points, the
(eps point, frame block) items of every n dealt round-robin over the ranks (dist.shard_items; the reference runs one OS
process per point and merges CSVs with tools/combine_data.py:64-95), one all-reduce of the counter block per n.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_dist.py [frames_scale]

Prints one JSON line per n and one for the whole sweep (rank 0): frames/s, edge updates/s, max-over-ranks device time;
with --check (1 extra single-GPU run per n on rank 0) asserts the N-GPU counters equal the 1-GPU counters bit for bit."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402
import iib_project_ldpc_codes_b200 as L  # noqa: E402
from iib_project_ldpc_codes_b200 import dist as D  # noqa: E402

os.environ.setdefault("LDPC_B200_CACHE", "0")
scale = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 1.0
check = "--check" in sys.argv
rank, world, local = D.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
eps = [round(0.34 + 0.11 * i / 19, 6) for i in range(20)]
cap, seed = 200, 12345
tot_frames = tot_eu = 0
tot_ms = 0.0
for n in (500, 1000, 2000, 4000, 8000):
    frames = int(scale * 4_000_000_000 // (20 * n)) // 128 * 128      # same number of edge updates per n
    vl, cl = L.generate_graph(n, 3, 6, seed=1, mode=1)
    g = L.Graph(vl, cl, device=local)
    items = D.shard_items(len(eps), frames, rank, world)
    d = torch.zeros((len(eps), cap + 6), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run(sd):
        d.zero_()
        for p, lo, cnt in items:
            g.mc_launch_point(float(eps[p]), p, sd, cnt, cap, d[p].data_ptr(), st, first_frame=lo)
        D.allreduce_counters(d)

    run(seed + 1)                      # warm: pilots, placement search (background), module load
    torch.cuda.synchronize(); g.wait_placement(); run(seed + 2)
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(seed); e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
    c = d.cpu().numpy().view(np.uint64)
    fr, ex = int(c[:, cap + 3].sum()), int(c[:, cap + 5].sum())
    same = None
    if check and rank == 0 and world > 1:
        d1 = torch.zeros_like(d)
        g.mc_launch(eps, seed, frames, cap, d1.data_ptr(), st)
        torch.cuda.synchronize()
        same = bool(torch.equal(d1, d))
        assert same, f"n={n}: the {world}-GPU counter block differs from the single-GPU block"
    if rank == 0:
        print(json.dumps({"n": n, "points": len(eps), "frames_per_point": frames, "n_gpus": world, "items_rank0": len(items),
                          "ms": float(ms.item()), "frames_per_s": fr / float(ms.item()) * 1e3,
                          "edge_updates_per_s": ex * n * 3 / float(ms.item()) * 1e3, "fer": (c[:, cap + 1] / np.maximum(c[:, cap + 3], 1)).round(6).tolist(),
                          "equal_to_single_gpu": same}), flush=True)
    tot_frames += fr; tot_eu += ex * n * 3; tot_ms += float(ms.item())
    g.close()
if rank == 0:
    print(json.dumps({"workload": "c5: (3,6) n in {500,1000,2000,4000,8000} x 20 eps points 0.34-0.45, (point, block) items over the ranks",
                      "n_gpus": world, "ms_total": tot_ms, "frames_per_s": tot_frames / tot_ms * 1e3,
                      "edge_updates_per_s": tot_eu / tot_ms * 1e3}), flush=True)
if world > 1:
    tdist.destroy_process_group()

"""Multi-GPU sharding of the Monte-Carlo frame loop: one process per GPU, no data-path collective.

The reference's parallelism is N independent OS processes whose CSVs are re-weighted and summed
afterwards (tools/combine_data.py:64-95).  Here frames are keyed by their GLOBAL index (the erasure
pattern of frame f depends only on (seed, point, f)), every rank decodes a disjoint, 128-aligned frame
range of every erasure point, and the only exchange is ONE all-reduce(sum) of the integer counter block
[npoints][iterations+6] (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).  Because the
counters are integers the result is bit-identical for any world size.
"""
from __future__ import annotations

import os

ALIGN = 128  # frames; ldpc_b200_mc_launch requires first_frame % 128 == 0


def shard_frames(num_frames: int, rank: int, world: int, first_frame: int = 0):
    """Contiguous 128-aligned split of [first_frame, first_frame+num_frames) -> (first, count) for `rank`.

    Every frame belongs to exactly one rank; ranks may get 0 frames when num_frames is small."""
    if first_frame % ALIGN:
        raise ValueError("first_frame must be a multiple of 128")
    blocks = (num_frames + ALIGN - 1) // ALIGN
    lo_b = blocks * rank // world
    hi_b = blocks * (rank + 1) // world
    lo = min(lo_b * ALIGN, num_frames)
    hi = min(hi_b * ALIGN, num_frames)
    return first_frame + lo, hi - lo


def shard_items(npoints: int, num_frames: int, rank: int, world: int, first_frame: int = 0, blocks_per_point: int = 0):
    """(point, block) partitioning -- SURVEY.md 8(e) "rank r takes (p, g) pairs", north_star "erasure-probability
    points shard across the GPUs"; the reference's analogue is one OS process per (eps, code number) whose CSVs
    tools/combine_data.py:64-95 re-weights and sums.

    Every point's frame range is cut into `blocks_per_point` 128-aligned blocks (default: the smallest count that
    gives every rank at least two items and makes npoints*blocks a multiple of `world`); item (b, p) has index
    b*npoints + p and goes to rank index % world.  Returns this rank's [(point, first_frame, count), ...] with
    consecutive blocks of a point merged.  Every (point, frame) belongs to exactly one rank, so the summed counter
    blocks are bit-identical to a single-GPU run whatever the partitioning."""
    if first_frame % ALIGN:
        raise ValueError("first_frame must be a multiple of 128")
    if blocks_per_point <= 0:
        blocks_per_point = 1
        while (npoints * blocks_per_point) % world or npoints * blocks_per_point < 2 * world:
            blocks_per_point += 1
    units = (num_frames + ALIGN - 1) // ALIGN
    mine = []
    for b in range(blocks_per_point):
        for p in range(npoints):
            if (b * npoints + p) % world != rank:
                continue
            lo = min(units * b // blocks_per_point * ALIGN, num_frames)
            hi = min(units * (b + 1) // blocks_per_point * ALIGN, num_frames)
            if hi > lo:
                mine.append((p, first_frame + lo, hi - lo))
    mine.sort()
    merged = []
    for p, lo, cnt in mine:
        if merged and merged[-1][0] == p and merged[-1][1] + merged[-1][2] == lo:
            merged[-1] = (p, merged[-1][1], merged[-1][2] + cnt)
        else:
            merged.append((p, lo, cnt))
    return merged


def init_from_env(backend: str | None = None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_counters(t):
    """In-place sum of the counter tensor (int64 view of the uint64 block) over all ranks."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def sharded_mc_run(graph, eps, seed, num_frames, iterations, expurgation=-1, first_frame=0, rank=0, world=1,
                   partition="frames"):
    """Decode this rank's share on its GPU and all-reduce; returns the GLOBAL uint64 counter block on the host.

    partition "frames": a contiguous slice of EVERY point's frames (perfectly balanced: erasure points near the
    threshold cost twice the iterations of points below it, and every rank gets the same mix) -- one launch;
    partition "points": (point, block) items dealt round-robin (shard_items) -- one launch per item, each on a single
    erasure point.  Both give bit-identical results."""
    import numpy as np
    import torch

    eps = np.atleast_1d(np.asarray(eps, dtype=np.float64))
    npoints = len(eps)
    dev = torch.device("cuda", graph.device)
    d = torch.zeros((npoints, iterations + 6), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        if partition == "frames":
            lo, cnt = shard_frames(num_frames, rank, world, first_frame)
            if cnt > 0:
                graph.mc_launch(eps, seed, cnt, iterations, d.data_ptr(), stream, expurgation=expurgation, first_frame=lo)
        elif partition == "points":
            for p, lo, cnt in shard_items(npoints, num_frames, rank, world, first_frame):
                # the pattern of a frame is keyed by its point index: launch point p as point p
                graph.mc_launch_point(float(eps[p]), p, seed, cnt, iterations, d[p].data_ptr(), stream,
                                      expurgation=expurgation, first_frame=lo)
        else:
            raise ValueError("partition must be 'frames' or 'points'")
        allreduce_counters(d)
        out = d.cpu().numpy().view(np.uint64)
    return out

#!/usr/bin/env python
"""bench.py -- throughput of the BEC message-passing Monte-Carlo hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c2|c1|c4|c5] [--impl ours|reference]

One "step" = one pass of the fused Monte-Carlo path over one batch of synthetic input: every erasure
point of the workload, `frames` frames per point PER GPU (weak scaling: rank r decodes the global frame
range [r*frames, (r+1)*frames) of every point; the only exchange is one all-reduce of the counter block).

Default workload = BASELINE.json configs[1]: (3,6)-regular n=10^4, eps in {0.38..0.43}, 10^7 frames per
point, iteration cap 200.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, dv, dc, eps list, frames per point per GPU, iteration cap, expurgation, description)
    "c2": dict(n=10000, dv=3, dc=6, eps=[0.38, 0.39, 0.40, 0.41, 0.42, 0.43], frames=10_000_000, cap=200,
               expurgation=-1, desc="(3,6)-regular n=10^4 BEC eps sweep 0.38-0.43, 10^7 frames per point, cap 200"),
    "c1": dict(n=1000, dv=3, dc=6, eps=[0.40], frames=100_000_000, cap=200, expurgation=-1,
               desc="(3,6)-regular n=1000 BEC eps=0.40, 10^8 frames (steady state of config 1), cap 200"),
    "c4": dict(n=2000, dv=4, dc=8, eps=[0.30, 0.32, 0.34, 0.36], frames=25_000_000, cap=200, expurgation=0,
               desc="(4,8)-regular n=2000 fixed graph, expurgation 0, eps 0.30-0.36, 2.5*10^7 frames per point"),
    "c3": dict(n=1_000_000, dv=3, dc=6, eps=[0.40, 0.42], frames=16384, cap=200, expurgation=-1,
               desc="(3,6)-regular n=10^6 BEC eps 0.40/0.42, graph streamed from HBM (engine E2), 16384 frames per point, cap 200"),
    "c4e": dict(n=2000, dv=4, dc=8, eps=[0.30, 0.32, 0.34, 0.36], frames=128 * 131072, cap=200, expurgation=0,
                ensemble=True, frames_per_graph=128,
                desc="expurgated (4,8)-regular ENSEMBLE n=2000 (parallel_simulator_expurgated.py run_simulation): fresh graph "
                     "per 128 frames generated in the kernel, expurgation 0, eps 0.30-0.36, 131072 graphs per point"),
    "c5": dict(n=4000, dv=3, dc=6, eps=[round(0.34 + 0.11 * i / 19, 6) for i in range(20)], frames=1_000_000,
               cap=200, expurgation=-1, desc="(3,6)-regular n=4000, 20 eps points 0.34-0.45, 10^6 frames per point"),
}
GRAPH_SEED = 1
RUN_SEED = 12345


def alg_bytes_per_edge_update(dv, dc, n, words=4):
    """state bits (2E + m + 2n per frame-iteration) + index bytes amortised over W=128 frames."""
    idx = 2 if n < 65536 else 4
    return (2 + 1 / dc + 2 / dv) / 8 + 2 * idx / 128


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref = the unmodified
    message_passing.c compiled here; falls back to the pinned C restatement) on all host threads, on a
    bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    import iib_project_ldpc_codes_b200 as L
    O.build()
    kind = "reference" if O.have_ref() else "port"
    chk = O.best()
    cores = os.cpu_count() or 1
    vl, cl = L.generate_graph(wl["n"], wl["dv"], wl["dc"], seed=GRAPH_SEED, mode=1)
    m = wl["n"] * wl["dv"] // wl["dc"]
    # bounded sample: ~64 frames per thread per point at n=10^4 scale
    fpp = args.ref_frames or max(cores * max(1, int(64 * 10000 / wl["n"])), 1 if wl["n"] > 100000 else 32)
    eps = wl["eps"]
    secs, frames, execs = [], 0, 0
    for s in range(args.warmup + args.steps):
        c, dsec = chk.mc_run(vl, cl, wl["n"], wl["n"] - m, wl["dv"], wl["dc"], eps, RUN_SEED, s * fpp * 0, fpp,
                             wl["cap"], wl["expurgation"], nthreads=cores, return_seconds=True)
        if s >= args.warmup:
            secs.append(dsec)
            frames += int(c[:, wl["cap"] + 3].sum())
            execs += int(c[:, wl["cap"] + 5].sum())
    total = sum(secs)
    value = frames / total
    eu = execs * wl["n"] * wl["dv"] / total
    sample = f"{fpp} frames/point x {len(eps)} points per step, decode calls only, {cores} threads"
    if wl.get("ensemble"):
        sample += " (one fixed graph: the reference's per-frame graph generation, 2-3 s per (4,8) n=2000 graph, is NOT timed)"
    line = {"impl": "reference", "metric": "decoded frames/s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "edge_updates_per_s": eu,
            "config": {"workload": args.workload + ": " + wl["desc"], "graph_seed": GRAPH_SEED, "run_seed": RUN_SEED},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample,
                             "edge_updates_per_s": eu},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def run_ensemble(args, wl, L, D, torch, tdist, rank, world, local, dev):
    """Ensemble workloads (a fresh random graph per frame group, generated inside the kernel): the public entry
    point is the synchronous host-buffer call ldpc_b200_mc_run_ensemble, so `value` and `e2e` are the same
    measurement (host eps in, host counter block out, every step)."""
    n, dv, dc, cap, eps = wl["n"], wl["dv"], wl["dc"], wl["cap"], wl["eps"]
    fpg = wl["frames_per_graph"]
    graphs = max(1, wl["frames"] // fpg)
    npoints = len(eps)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def step(s):
        c, st = L.mc_run_ensemble(n, dv, dc, eps, RUN_SEED + s, graphs, cap, frames_per_graph=fpg, graph_seed=GRAPH_SEED,
                                  gen_mode=2, first_graph=(s * world + rank) * graphs, expurgation=wl["expurgation"],
                                  device=local, return_stats=True)
        t = torch.from_numpy(c.view(np.int64).copy()).to(dev)
        D.allreduce_counters(t)
        return t, st

    for s in range(args.warmup):
        step(s)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    acc, stats = None, None
    phys = 0
    for s in range(args.steps):
        t, stats = step(args.warmup + s)
        acc = t if acc is None else acc + t
        phys += stats["phys_frame_iterations"]
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank != 0:
        return
    cnt = acc.cpu().numpy().view(np.uint64)
    tot_frames, tot_exec = int(cnt[:, cap + 3].sum()), int(cnt[:, cap + 5].sum())
    E = n * dv
    value = tot_frames / (ms * 1e-3)
    b_alg = alg_bytes_per_edge_update(dv, dc, n)
    smem_peak = L.api.smem_bandwidth_gbs(local)
    achieved = tot_exec / world * E * b_alg / (ms * 1e-3) / 1e9
    line = {"metric": "decoded frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "edge_updates_per_s": tot_exec * E / (ms * 1e-3),
            "graphs_per_s": npoints * graphs * world * args.steps / (ms * 1e-3),
            "config": {"workload": args.workload + ": " + wl["desc"], "graphs_per_point_per_gpu": graphs,
                       "frames_per_graph": fpg, "graph_seed": GRAPH_SEED, "run_seed": RUN_SEED,
                       "l2": "working set is on-chip (graph, state and generator scratch in shared memory)", "engine": stats},
            "roofline": {"bound": "smem", "achieved": achieved, "peak": smem_peak, "unit": "GB/s", "frac": achieved / smem_peak,
                         "traffic": None, "kernel": "e1_mc_kernel (ensemble: graph generation + decode), whole call timed",
                         "alg_bytes_per_edge_update": b_alg,
                         "lane_inflation": phys * world / max(tot_exec, 1),
                         "note": "graph generation and the expurgation second pass are inside the timed call and carry no "
                                 "algorithmic bytes"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": int(npoints * 8),
                    "d2h_bytes_per_step": int(npoints * (cap + 6) * 8)},
            "gpu_launches": int(stats["kernel_launches"]) * args.steps, "clocks": clocks}
    _emit(line)
    if world > 1:
        tdist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; keep stdout for the ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per point per GPU per step (default: workload's)")
    ap.add_argument("--ref-frames", type=int, default=0, help="reference arm: frames per point per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra-c1", action="store_true", help="also report n=1000 steady-state frames/s")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.frames:
        wl["frames"] = args.frames
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import iib_project_ldpc_codes_b200 as L
    from iib_project_ldpc_codes_b200 import dist as D

    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as tdist
    if wl.get("ensemble"):
        return run_ensemble(args, wl, L, D, torch, tdist, rank, world, local, dev)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    n, dv, dc, cap, eps = wl["n"], wl["dv"], wl["dc"], wl["cap"], wl["eps"]
    vl, cl = L.generate_graph(n, dv, dc, seed=GRAPH_SEED, mode=1)
    g = L.Graph(vl, cl, device=local)
    # the handle is reused for every step: let the one-time bank-conflict placement search run longer (public knob)
    place_effort = max(6, min(24, 24000 // n))
    g.set_place_effort(place_effort)
    frames = wl["frames"]
    npoints = len(eps)
    d_cnt = torch.zeros((npoints, cap + 6), dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step(s):
        flush.fill_(s & 0xFF)  # L2 flush between steps (the kernel's working set is on-chip anyway)
        d_cnt.zero_()
        g.mc_launch(eps, RUN_SEED + s, frames, cap, d_cnt.data_ptr(), stream, expurgation=wl["expurgation"],
                    first_frame=rank * ((frames + 127) // 128 * 128))
        D.allreduce_counters(d_cnt)

    for s in range(args.warmup):
        step(s)
    barrier()
    g.kernel_times_ms()
    g.mc_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_frames = tot_exec = 0
    barrier()
    e0.record()
    for s in range(args.steps):
        step(args.warmup + s)
        # counters of the step are consumed after the loop for the last step; accumulate on device
        if s == 0:
            acc = d_cnt.clone()
        else:
            acc += d_cnt
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
    ms = float(ms.item())
    cnt = acc.cpu().numpy().view(np.uint64)
    tot_frames = int(cnt[:, cap + 3].sum())
    tot_exec = int(cnt[:, cap + 5].sum())
    kt = g.kernel_times_ms()
    stats = g.mc_stats()
    E = n * dv
    value = tot_frames / (ms * 1e-3)
    eu_per_s = tot_exec * E / (ms * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launches; algorithmic bytes / measured duration)
    b_alg = alg_bytes_per_edge_update(dv, dc, n)
    per_rank_exec = tot_exec / world
    k_ms = float(np.mean(kt)) if len(kt) else float("nan")
    alg_bytes_per_launch = per_rank_exec / args.steps * E * b_alg
    achieved = alg_bytes_per_launch / (k_ms * 1e-3) / 1e9
    smem_peak = L.api.smem_bandwidth_gbs(local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    e1 = stats["engine"] == 1
    roofline = {"bound": "smem" if e1 else "hbm", "achieved": achieved, "peak": smem_peak if e1 else hbm_peak,
                "unit": "GB/s", "frac": achieved / (smem_peak if e1 else hbm_peak), "traffic": None,
                "note": ("E1 keeps graph+state in shared memory: HBM sees only the counter REDs and the spill records of the "
                         "staged pipeline (ncu, profiles/r1j_e1_staged_c2_w64_stage0_ncu_full.txt: 1.7 MB read + 0.51 GB written "
                         "by the stage-0 launch of 1.2*10^6 frames, 5*10^-4 of its algorithmic shared-memory bytes); peak = "
                         "shared-memory read bandwidth measured on this GPU by ldpc_b200_smem_probe (conflict-free LDS.128 "
                         "from all SMs)")
                if e1 else "E2 streams frame-minor state rows and int32 lookups from HBM/L2; peak = measured HBM copy bandwidth",
                "kernel": (("e1_mc_staged_kernel x%d stages (survivor re-packing), timed as one pipeline" % stats["stages"])
                           if stats["stages"] > 1 else "e1_mc_kernel") if e1 else "e2_kernel",
                "kernel_ms_mean": k_ms,
                "kernel_share_of_step": k_ms * args.steps / ms,
                "alg_bytes_per_edge_update": b_alg, "alg_bytes_per_launch": alg_bytes_per_launch,
                "frac_of_hbm_peak": achieved / hbm_peak, "hbm_peak": hbm_peak,
                "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                "lane_inflation": stats["phys_frame_iterations"] / max(per_rank_exec, 1)}

    # ---- end to end through the public host-buffer API (graph upload + eps H2D + counters D2H per step)
    e2e_frames = 0
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        # the call a user makes per batch: host eps in, host counter block out, on an existing graph handle
        c = g.mc_run(eps, RUN_SEED + 1000 + s, frames, cap, expurgation=wl["expurgation"],
                     first_frame=rank * ((frames + 127) // 128 * 128))
        e2e_frames += int(c[:, cap + 3].sum())
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t_e2e, op=tdist.ReduceOp.MAX)
    e2e = {"value": e2e_frames * world / float(t_e2e.item()), "unit": "frames/s",
           "h2d_bytes_per_step": int(npoints * 16),
           "d2h_bytes_per_step": int(npoints * (cap + 6) * 8)}

    if rank != 0:
        return
    line = {"metric": "decoded frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "edge_updates_per_s": eu_per_s,
            "config": {"workload": args.workload + ": " + wl["desc"], "frames_per_point_per_gpu": frames,
                       "graph_seed": GRAPH_SEED, "run_seed": RUN_SEED, "l2": "256 MiB flush write between steps",
                       "place_effort": place_effort, "engine": stats},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(stats["kernel_launches"]), "clocks": clocks}

    if not args.no_cpu_baseline and world >= 1:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        chk = O.best()
        fpp = max(cores * max(1, int(64 * 10000 / n)), 1 if n > 100000 else 32)
        c, dsec = chk.mc_run(vl, cl, n, g.k, dv, dc, eps, RUN_SEED, 0, fpp, cap, wl["expurgation"], nthreads=cores,
                             return_seconds=True)
        v = int(c[:, cap + 3].sum()) / dsec
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores,
                                "kind": "reference" if O.have_ref() else "port",
                                "sample": f"{fpp} frames/point x {npoints} points, decode calls only",
                                "edge_updates_per_s": int(c[:, cap + 5].sum()) * E / dsec}
        # the sample is also a live parity check of the GPU path on the same frames
        got = g.mc_run(eps, RUN_SEED, fpp, cap, expurgation=wl["expurgation"])
        line["cpu_baseline"]["parity_on_sample"] = bool(np.array_equal(got, c))

    if args.extra_c1 or args.workload == "c2":
        w1 = WORKLOADS["c1"]
        v1, c1 = L.generate_graph(w1["n"], 3, 6, seed=GRAPH_SEED, mode=1)
        g1 = L.Graph(v1, c1, device=local)
        g1.set_place_effort(24)
        f1 = 20_000_000
        d1 = torch.zeros((1, w1["cap"] + 6), dtype=torch.int64, device=dev)
        for _ in range(2):
            g1.mc_launch(w1["eps"], RUN_SEED, f1, w1["cap"], d1.data_ptr(), stream)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d1.zero_()
        a0.record()
        g1.mc_launch(w1["eps"], RUN_SEED + 1, f1, w1["cap"], d1.data_ptr(), stream)
        a1.record()
        torch.cuda.synchronize()
        t1 = a0.elapsed_time(a1) * 1e-3
        c1n = d1.cpu().numpy().view(np.uint64)
        line["extra_n1000"] = {"workload": w1["desc"].replace("10^8", "2*10^7"), "frames_per_s_per_gpu": f1 / t1,
                               "edge_updates_per_s_per_gpu": int(c1n[0, w1["cap"] + 5]) * 3000 / t1,
                               "roofline_frac_smem": int(c1n[0, w1["cap"] + 5]) * 3000 / t1 *
                               alg_bytes_per_edge_update(3, 6, 1000) / 1e9 / smem_peak,
                               "engine": g1.mc_stats()}
    _emit(line)
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""device-timed mc_launch vs wall-clock mc_run on one handle, alternating (why do they differ?)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iib_project_ldpc_codes_b200 as L
n = int(sys.argv[1]); npts = int(sys.argv[2]); frames = int(sys.argv[3]); flush_on = int(sys.argv[4]) if len(sys.argv) > 4 else 1
eps = [round(0.34 + 0.11 * i / max(npts - 1, 1), 6) for i in range(npts)] if npts > 1 else [0.40]
cap = 200
vl, cl = L.generate_graph(n, 3, 6, seed=1, mode=1)
g = L.Graph(vl, cl)
dev = torch.device("cuda", 0)
d = torch.zeros((len(eps), cap + 6), dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    g.mc_launch(eps, rep, frames, cap, d.data_ptr(), st); torch.cuda.synchronize()
g.kernel_times_ms()
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    if flush_on: flush.fill_(rep)
    d.zero_()
    t0 = time.perf_counter()
    g.mc_launch(eps, 10 + rep, frames, cap, d.data_ptr(), st)
    t1 = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    kt = g.kernel_times_ms()
    t2 = time.perf_counter()
    c = g.mc_run(eps, 20 + rep, frames, cap)
    t3 = time.perf_counter()
    print(f"launch: events {e0.elapsed_time(e1):.2f} ms (kernel bracket {sum(kt):.2f} ms, host call {1e3*(t1-t0):.2f} ms)   mc_run wall {1e3*(t3-t2):.2f} ms", flush=True)

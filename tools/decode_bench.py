import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
import iib_project_ldpc_codes_b200 as L
for n, frames in [(1000, 200000), (10000, 20000)]:
    vl, cl = L.generate_graph(n,3,6,seed=1,mode=1)
    g = L.Graph(vl, cl)
    pats = L.draw_patterns(n, 0.40, 1, 0, 0, min(frames, 65535*32))
    g.decode_batch(pats[:1000], 200)
    t=time.time(); M,e,r = g.decode_batch(pats, 200); dt=time.time()-t
    print(f"decode_batch n={n} frames={frames}: {frames/dt:.3g} frames/s (host buffers, cap 200), mean ret {r.mean():.2f}")

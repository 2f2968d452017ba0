import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
import iib_project_ldpc_codes_b200 as L
n,dv,dc=1000,3,6
eps=[0.40]
for fpg, graphs in ((1024, 32768),)*10:
        t=time.time(); c,st=L.mc_run_ensemble(n,dv,dc,eps,1,graphs,200,frames_per_graph=fpg,expurgation=-1,gen_mode=2,return_stats=True); dt=time.time()-t
        fr=int(c[0,203]); ex=int(c[0,205])
        print(f"fpg={fpg} graphs={graphs}: {dt:.3f} s  {fr/dt:.3g} frames/s frames={fr} launches={st['kernel_launches']} grid={st['grid']} FER={c[0,201]/fr:.4f}", flush=True)

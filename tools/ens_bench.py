import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
import iib_project_ldpc_codes_b200 as L
for (n,dv,dc,eps,exp) in [(2000,4,8,[0.34],0),(1000,3,6,[0.40],-1)]:
    for mode, graphs in ((1, 65536), (2, 65536), (2, 524288)):
        L.mc_run_ensemble(n,dv,dc,eps,1,256,200,frames_per_graph=128,expurgation=exp,gen_mode=mode)
        t=time.time(); c,st=L.mc_run_ensemble(n,dv,dc,eps,1,graphs,200,frames_per_graph=128,expurgation=exp,gen_mode=mode,return_stats=True); dt=time.time()-t
        print(f"n={n} ({dv},{dc}) eps={eps} gen_mode={mode} graphs={graphs}: {dt:.3f} s  {graphs*128/dt:.3g} frames/s  {graphs/dt:.3g} graphs/s  FER={c[0,201]/c[0,203]:.4g} teams={st['teams']}x{st['team_warps']} words={st['words']}")

import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
import iib_project_ldpc_codes_b200 as L
for (n,dv,dc,eps,exp) in [(2000,4,8,[0.34],0),(2000,4,8,[0.34],-1),(1000,3,6,[0.40],-1)]:
    for fpg in (128, 256, 1024, 4096):
        graphs = (1 << 25) // fpg
        L.mc_run_ensemble(n,dv,dc,eps,1,256,200,frames_per_graph=fpg,expurgation=exp,gen_mode=2)
        t=time.time(); c,st=L.mc_run_ensemble(n,dv,dc,eps,1,graphs,200,frames_per_graph=fpg,expurgation=exp,gen_mode=2,return_stats=True); dt=time.time()-t
        fr=int(c[0,203]); ex=int(c[0,205])
        print(f"n={n} ({dv},{dc}) exp={exp} fpg={fpg} graphs={graphs}: {dt:.3f} s  {fr/dt:.3g} frames/s  {graphs/dt:.3g} graphs/s  it/frame={ex/fr:.2f} infl={st['phys_frame_iterations']/ex:.2f} second={st['second_pass_groups']} teams={st['teams']}x{st['team_warps']} words={st['words']}", flush=True)

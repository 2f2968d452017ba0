import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
import iib_project_ldpc_codes_b200 as L
n,dv,dc,eps=2000,4,8,[0.34]
L.mc_run_ensemble(n,dv,dc,eps,1,256,200,frames_per_graph=128,expurgation=0)
for graphs in (256, 2048, 8192, 32768, 131072):
    t=time.time(); c,st=L.mc_run_ensemble(n,dv,dc,eps,1,graphs,200,frames_per_graph=128,expurgation=0,return_stats=True); dt=time.time()-t
    print('graphs',graphs,'time %.3f s'%dt,'frames/s %.3g'%(graphs*128/dt),'graphs/s %.3g'%(graphs/dt))

for w in 4 2; do
LDPC_B200_WORDS=$w python - <<'PY'
import sys, time, os
sys.path.insert(0,'.')
import iib_project_ldpc_codes_b200 as L
for (n,dv,dc,eps,exp) in [(2000,4,8,[0.30,0.32,0.34,0.36],0),(1000,3,6,[0.40],-1)]:
    L.mc_run_ensemble(n,dv,dc,eps,1,256,200,frames_per_graph=128,expurgation=exp,gen_mode=2)
    for rep in range(2):
        t=time.time(); c,st=L.mc_run_ensemble(n,dv,dc,eps,1,65536,200,frames_per_graph=128,expurgation=exp,gen_mode=2,return_stats=True); dt=time.time()-t
        fr=int(c[:,203].sum()); ex=int(c[:,205].sum())
        print(f"WORDS={os.environ['LDPC_B200_WORDS']} n={n} ({dv},{dc}) exp={exp}: {dt:.3f} s  {fr/dt:.3g} frames/s infl={st['phys_frame_iterations']/ex:.2f} second={st['second_pass_groups']} teams={st['teams']}x{st['team_warps']} words={st['words']} smem={st['smem_bytes']}", flush=True)
PY
done

for t in 2 0; do export LDPC_B200_E2_TMA=$t; echo "E2_TMA=$t"; python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-legs --no-cold 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('c3 value %.4g EU/s %.4g frac %.3f ms/step %.1f infl %.3f'%(d['value'], d['edge_updates_per_s'], r['frac'], d['ms_per_step'], r['lane_inflation']))"; done
unset LDPC_B200_E2_TMA
python tools/c5_dist.py 32 > gpurun_out/r2e_c5_points_scale32_1gpu.jsonl 2>/dev/null; tail -1 gpurun_out/r2e_c5_points_scale32_1gpu.jsonl

timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; tail -5 gpurun_out/r2d_tests.log
for t in 2 0; do export LDPC_B200_E2_TMA=$t; echo "E2_TMA=$t"; python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-legs --no-cold 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('c3 value %.4g EU/s %.4g frac %.3f ms/step %.1f infl %.3f'%(d['value'], d['edge_updates_per_s'], r['frac'], d['ms_per_step'], r['lane_inflation']))"; done
unset LDPC_B200_E2_TMA
python tools/dropin_bench.py 1000 3
python tools/dropin_bench.py 10000 2

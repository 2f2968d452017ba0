timeout 900 python -m pytest tests/test_gpu_e2.py -x -q > gpurun_out/e2_tests.log 2>&1; tail -4 gpurun_out/e2_tests.log
for d in 0 2 3 4; do export LDPC_B200_E2_ASYNC=$d; echo "E2_ASYNC=$d"; python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-legs --no-cold 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('c3 value %.4g EU/s %.4g frac %.3f ms/step %.1f grid %d'%(d['value'], d['edge_updates_per_s'], r['frac'], d['ms_per_step'], d['engine']['grid']))"; done

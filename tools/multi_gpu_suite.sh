# multi-GPU measurement suite.  usage: bash tools/multi_gpu_suite.sh N TAG [full]
N=${1:-2}; TAG=${2:-r2x}; FULL=${3:-}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
set -x
$TR bench.py --gpus $N --scaling strong --steps 3 --warmup 3 --no-cpu-baseline --no-legs --no-cold > $O/${TAG}_c2_strong_${N}gpu.json 2> $O/${TAG}_c2_strong_${N}gpu.err
if [ -n "$FULL" ]; then
$TR bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --no-legs --no-cold > $O/${TAG}_c2_weak_${N}gpu.json 2> $O/${TAG}_c2_weak_${N}gpu.err
$TR bench.py --gpus $N --workload c4e --steps 2 --warmup 2 --no-cpu-baseline --no-legs --no-cold > $O/${TAG}_c4e_weak_${N}gpu.json 2> $O/${TAG}_c4e_weak_${N}gpu.err
$TR bench.py --gpus $N --workload c1 --steps 3 --warmup 3 --no-cpu-baseline --no-legs --no-cold > $O/${TAG}_c1_weak_${N}gpu.json 2> $O/${TAG}_c1_weak_${N}gpu.err
$TR tools/c5_dist.py 1.0 --check > $O/${TAG}_c5_points_${N}gpu.jsonl 2> $O/${TAG}_c5_points_${N}gpu.err
fi
for f in $O/${TAG}_*_${N}gpu.json; do python - <<PY
import json
try:
    d=json.load(open("$f")); print("$f", "n_gpus", d["n_gpus"], "%.4g frames/s"%d["value"], "scaling", d["scaling"], "ms/step %.1f"%d["ms_per_step"], d.get("strong_scaling_counters_equal_single_gpu"))
except Exception as e: print("$f failed", e)
PY
done
tail -2 $O/${TAG}_c5_points_${N}gpu.jsonl 2>/dev/null

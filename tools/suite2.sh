# round-2 measurement suite (short): tests, bench lines, launch list, ncu captures.  usage: bash tools/suite2.sh TAG
TAG=${1:-r2x}
O=gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log
python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
for w in c1 c4 c4e c5; do python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-legs --no-cold > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_bench_c2_launches_raw.csv python bench.py --steps 2 --warmup 1 --frames 2000000 --no-cpu-baseline --no-legs --no-cold > $O/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:e1_mc_staged --launch-skip 1 -c 1 -f -o $O/prof_c1_${TAG} python tools/sweep.py 1000 3 6 0.40 2000000 > $O/${TAG}_ncu_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:e1_mc_staged --launch-skip 1 -c 1 -f -o $O/prof_c2_${TAG} python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 200000 > $O/${TAG}_ncu_c2.log 2>&1
head -c 1500 $O/${TAG}_bench_c2.json; echo
for w in c1 c4 c4e c5; do python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$w.json"))
    print("$w", "%.4g frames/s"%d["value"], "%.4g EU/s"%d["edge_updates_per_s"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e: print("$w", "failed", e)
PY
done

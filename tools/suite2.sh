# round-2 measurement suite: tests, bench lines, launch list, ncu captures.  usage: bash tools/suite2.sh TAG
TAG=${1:-r2x}
O=gpurun_out
set -x
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log; fi
python bench.py --impl reference > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err
python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
for w in c1 c4 c4e c5; do python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-legs --no-cold > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; done
python tools/dropin_bench.py 1000 3 > $O/${TAG}_dropin.log 2>&1
python tools/n_sweep.py > $O/${TAG}_c5_n_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_bench_c2_launches_raw.csv python bench.py --steps 2 --warmup 1 --frames 2000000 --no-cpu-baseline --no-legs --no-cold > $O/${TAG}_ncu_launches.log 2>&1
LDPC_B200_PLACE_SYNC=1 ncu --set full --clock-control none --import-source on -k regex:e1_mc_staged --launch-skip 1 -c 1 -f -o $O/prof_c1_${TAG} python tools/sweep.py 1000 3 6 0.40 2000000 > $O/${TAG}_ncu_c1.log 2>&1
LDPC_B200_PLACE_SYNC=1 ncu --set full --clock-control none --import-source on -k regex:e1_mc_overlap --launch-skip 0 -c 1 -f -o $O/prof_c2_${TAG} python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 200000 > $O/${TAG}_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:e2_kernel --launch-skip 1 -c 1 -f -o $O/prof_c3_${TAG} python tools/sweep.py 1000000 3 6 0.40,0.42 4096 > $O/${TAG}_ncu_c3.log 2>&1
# summarise the captures on the box (the reports are 20+ MB each: only the text comes back)
for k in c1 c2 c3; do
  if [ -f $O/prof_${k}_${TAG}.ncu-rep ]; then
    python tools/ncu_summary.py $O/prof_${k}_${TAG}.ncu-rep > $O/${TAG}_${k}_ncu_full.txt 2>&1
    python tools/ncu_regions.py $O/prof_${k}_${TAG}.ncu-rep >> $O/${TAG}_${k}_ncu_full.txt 2>&1
    rm -f $O/prof_${k}_${TAG}.ncu-rep
  fi
done
head -c 600 $O/${TAG}_bench_c2.json; echo
for w in c1 c4 c4e c5; do python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$w.json"))
    print("$w", "%.4g frames/s"%d["value"], "%.4g EU/s"%d["edge_updates_per_s"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e: print("$w", "failed", e)
PY
done

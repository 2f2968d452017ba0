LDPC_B200_PLACE_EFFORT=8 python tools/sweep.py 1000 3 6 0.40 20000000 2>&1 | cut -c1-330
LDPC_B200_PLACE_EFFORT=32 python tools/sweep.py 1000 3 6 0.40 20000000 2>&1 | cut -c1-330
python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200

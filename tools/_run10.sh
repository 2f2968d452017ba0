timeout 600 python -m pytest tests -m gpu -x -q -k "overflow" > gpurun_out/s12_tests.log 2>&1
tail -3 gpurun_out/s12_tests.log
python bench.py --workload c4e --steps 2 --warmup 1 > gpurun_out/bench_c4e.json 2> gpurun_out/bench_c4e.err
cat gpurun_out/bench_c4e.json; tail -3 gpurun_out/bench_c4e.err

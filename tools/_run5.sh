python tools/sweep.py 1000 3 6 0.40 20000000 > gpurun_out/s7_c1.log 2>&1
python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 > gpurun_out/s7_c2.log 2>&1
cat gpurun_out/s7_c1.log gpurun_out/s7_c2.log
timeout 600 python -m pytest tests -m gpu -x -q -k "parity or staged" > gpurun_out/s7_tests.log 2>&1
tail -3 gpurun_out/s7_tests.log

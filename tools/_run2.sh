set -x
python tools/sweep.py 1000 3 6 0.40 20000000 > gpurun_out/s2_c1.log 2>&1
python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 > gpurun_out/s2_c2.log 2>&1
cat gpurun_out/s2_c1.log gpurun_out/s2_c2.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1
tail -5 gpurun_out/s2_tests.log

timeout 900 python -m pytest tests -m gpu -x -q -k "ensemble or drivers" > gpurun_out/s11_tests.log 2>&1
tail -4 gpurun_out/s11_tests.log
python tools/ens_bench.py > gpurun_out/s11_ens.log 2>&1
cat gpurun_out/s11_ens.log

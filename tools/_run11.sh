python tools/ens_bench2.py > gpurun_out/s13_ens.log 2>&1
cat gpurun_out/s13_ens.log

./tools/ubench/smem_rate > gpurun_out/smem_rate.log 2>&1
cat gpurun_out/smem_rate.log
python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 > gpurun_out/s4_c2.log 2>&1
cat gpurun_out/s4_c2.log

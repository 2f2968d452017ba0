python tools/sweep.py 1000 3 6 0.40 20000000 4,8,3 4,6,4 4,8,2 4,7,3 > gpurun_out/s8_c1.log 2>&1
python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 2,1,24 2,1,16 2,1,20 > gpurun_out/s8_c2.log 2>&1
cat gpurun_out/s8_c1.log gpurun_out/s8_c2.log | cut -c1-150

set -x
python tools/sweep.py 1000 3 6 0.40 20000000 > gpurun_out/s1_c1.log 2>&1
python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 > gpurun_out/s1_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:e1_mc_staged --launch-skip 1 -c 1 -f -o gpurun_out/prof_c1_r1g python tools/sweep.py 1000 3 6 0.40 2000000 > gpurun_out/ncu_c1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:e1_mc_staged --launch-skip 1 -c 1 -f -o gpurun_out/prof_c2_r1g python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 200000 > gpurun_out/ncu_c2g.log 2>&1
cat gpurun_out/s1_c1.log gpurun_out/s1_c2.log

export LDPC_B200_LIB=$PWD/iib_project_ldpc_codes_b200/libvar_t1024.so
python tools/sweep.py 1000 3 6 0.40 20000000 4,8,4 4,8,3 4,7,4 > gpurun_out/s9_c1.log 2>&1
python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 2,1,32 2,1,28 > gpurun_out/s9_c2.log 2>&1
cat gpurun_out/s9_c1.log gpurun_out/s9_c2.log | cut -c1-150

LDPC_B200_TRACE=1 python tools/ens_bench3.py > gpurun_out/s14_ens.log 2>&1
cat gpurun_out/s14_ens.log

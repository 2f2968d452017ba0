for s in 6 8 10 12; do
  echo "STAGES=$s"
  LDPC_B200_STAGES=$s python tools/sweep.py 1000 3 6 0.40 20000000 2>&1 | cut -c1-200
  LDPC_B200_STAGES=$s python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 2>&1 | cut -c1-200
done > gpurun_out/s10.log 2>&1
cat gpurun_out/s10.log

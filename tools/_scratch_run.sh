timeout 900 python -m pytest tests/test_gpu_staged.py tests/test_gpu_fuzz.py -x -q > gpurun_out/ov_tests.log 2>&1; tail -4 gpurun_out/ov_tests.log
run() { timeout 300 python tools/sweep.py "$@" 2>&1 | tail -1 | sed -e 's/ place=.*overlap:/ overlap:/' -e 's/ place=.*//' | cut -c1-400; }
for l in 2 0; do export LDPC_B200_OVERLAP_LATER=$l; echo "c2 overlap_later=$l"; run 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 2000000; run 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 2000000; done
for l in 2 0; do export LDPC_B200_OVERLAP_LATER=$l; echo "n=8000 overlap_later=$l"; run 8000 3 6 0.38,0.40,0.42 3000000; done

set -x
./tools/ubench/smem_rate > gpurun_out/smem_rate.log 2>&1
cat gpurun_out/smem_rate.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1
tail -5 gpurun_out/s3_tests.log

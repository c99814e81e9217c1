set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pool" > gpurun_out/gpu_tests12.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests12.log
timeout 300 python scripts/time_pool_mma.py 0 0 rows > gpurun_out/time_pm12.log 2>&1
echo done

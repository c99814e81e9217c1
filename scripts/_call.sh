set -x
timeout 600 python scripts/dbg_pool_mma.py > gpurun_out/dbg_pool_mma2.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_pool_mma2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pool" > gpurun_out/gpu_tests4.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests4.log
echo done

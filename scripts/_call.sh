set -x
timeout 300 python scripts/time_pool_mma.py rows 0 0x30000 0 > gpurun_out/time_pm32.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_canaries.py -m gpu -q -x -k "pool or canar" > gpurun_out/gpu_tests32.log 2>&1
XM3D_SO=xmask3d_b200/libxm3d_dbg.so timeout 600 python scripts/exp_pool_mma.py 0 0x30000 > gpurun_out/exp_pm32.log 2>&1
echo done

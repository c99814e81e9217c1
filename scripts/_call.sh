set -x
XM3D_SO=xmask3d_b200/libxm3d_prev.so timeout 300 python scripts/time_pool_mma.py 0 0 > gpurun_out/time_pm46p.log 2>&1
timeout 300 python scripts/time_pool_mma.py 0 0 > gpurun_out/time_pm46n.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_canaries.py -m gpu -q -x -k "pool or canar" > gpurun_out/gpu_tests46.log 2>&1
echo done

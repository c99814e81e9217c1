set -x
XM3D_SO=xmask3d_b200/libxm3d_head.so timeout 300 python scripts/time_pool_mma.py 0 0 > gpurun_out/time_pm43h.log 2>&1
timeout 300 python scripts/time_pool_mma.py 0 0 > gpurun_out/time_pm43c.log 2>&1
XM3D_SO=xmask3d_b200/libxm3d_head.so timeout 300 python scripts/time_pool_mma.py 0 0 >> gpurun_out/time_pm43h.log 2>&1
timeout 300 python scripts/time_pool_mma.py 0 0 >> gpurun_out/time_pm43c.log 2>&1
echo done

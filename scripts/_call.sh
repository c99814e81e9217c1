set -x
timeout 600 python scripts/dbg_pool_mma.py > gpurun_out/dbg_pool_mma2.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_pool_mma2.log
XM3D_SO=xmask3d_b200/libxm3d_dbg.so timeout 600 python scripts/exp_pool_mma.py 0 > gpurun_out/exp_pool_mma2.log 2>&1; echo "rc=$?" >> gpurun_out/exp_pool_mma2.log
echo done

set -x
XM3D_SO=xmask3d_b200/libxm3d_dbg.so timeout 600 python scripts/exp_pool_mma.py 0 0x22 > gpurun_out/exp_pool_mma.log 2>&1; echo "rc=$?" >> gpurun_out/exp_pool_mma.log
echo done

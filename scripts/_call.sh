set -x
timeout 600 python scripts/dbg_pool_mma.py > gpurun_out/dbg_pool_mma2.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_pool_mma2.log
timeout 600 python scripts/time_pool_mma.py 0 0x400 > gpurun_out/time_pool_mma.log 2>&1; echo "rc=$?" >> gpurun_out/time_pool_mma.log
echo done

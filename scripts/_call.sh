set -x
timeout 300 python scripts/prof_pool_mma.py 0 > gpurun_out/prof36_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pool_mma2_kernel -c 1 -s 2 -f -o gpurun_out/prof_r02_pool_mma2_v2 python scripts/prof_pool_mma.py 0 > gpurun_out/prof36_ncu.log 2>&1
echo done

set -x
B="python bench.py --profile-steps 2 --no-cpu --extras none"
$B > gpurun_out/p_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/p_ncu1.log 2>&1
$B > gpurun_out/p_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pool_sum_kernel -s 1 -c 1 -o gpurun_out/prof_r02_pool_sum $B > gpurun_out/p_ncu2.log 2>&1
python scripts/prof_pool_mma.py 0 > gpurun_out/p_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pool_mma2 -s 2 -c 1 -o gpurun_out/prof_r02_pool_mma2 python scripts/prof_pool_mma.py 0 > gpurun_out/p_ncu3.log 2>&1
python scripts/prof_kernels.py > gpurun_out/p_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:point_logits_kernel -c 4 -o gpurun_out/prof_r02_point_logits python scripts/prof_kernels.py > gpurun_out/p_ncu4.log 2>&1
echo done

set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,temperature.gpu --format=csv > gpurun_out/time_pm18.log
timeout 600 python scripts/time_pool_mma.py rows 0 0x30000 0 rows >> gpurun_out/time_pm18.log 2>&1
echo done

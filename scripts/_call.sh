set -x
timeout 600 python scripts/exp_timeline.py > gpurun_out/timeline.log 2>&1; echo "rc=$?" >> gpurun_out/timeline.log
echo done

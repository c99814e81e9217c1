set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "point_logits" > gpurun_out/gpu_tests8.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests8.log
timeout 600 python scripts/time_point_logits.py > gpurun_out/time_point_logits.log 2>&1
echo done

set -x
timeout 600 python -m pytest tests/test_canaries.py -m gpu -q -x > gpurun_out/gpu_tests11.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests11.log
echo done

set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests3.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests3.log
echo done

set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "many_small" > gpurun_out/gpu_tests40.log 2>&1
echo done

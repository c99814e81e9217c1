set -x
timeout 900 python -m pytest tests/test_contra.py tests/test_gpu_parity.py -m gpu -x -q -k "contra or scene_mean or abi" > gpurun_out/gpu_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests2.log
echo done

set -x
timeout 900 python -m pytest tests -m gpu -q -k "logits or autograd or abi" > gpurun_out/gpu_tests5.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests5.log
timeout 300 python scripts/time_logits.py > gpurun_out/time_logits.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
echo done

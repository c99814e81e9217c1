set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests6.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
echo done

set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests49.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests49.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke49.log 2>&1; echo "rc=$?" >> gpurun_out/smoke49.log
timeout 900 python bench.py > gpurun_out/bench49.json 2> gpurun_out/bench49.err; echo "rc=$?" >> gpurun_out/bench49.err
echo done

set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests13.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests13.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke13.log 2>&1; echo "rc=$?" >> gpurun_out/smoke13.log
timeout 900 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "rc=$?" >> gpurun_out/bench13.err
timeout 900 python bench.py --impl reference > gpurun_out/bench13_ref.json 2> gpurun_out/bench13_ref.err; echo "rc=$?" >> gpurun_out/bench13_ref.err
echo done

set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests10.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests10.log
timeout 600 python bench.py --steps 20 --warmup 5 --masks overlap --extras none --no-cpu > gpurun_out/bench_r02_overlap.json 2> gpurun_out/bench_r02_overlap.err; echo "rc=$?" >> gpurun_out/bench_r02_overlap.err
echo done

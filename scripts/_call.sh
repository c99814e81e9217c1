set -x
timeout 600 python scripts/time_pool_mma.py 0 0x200 0x100 0x300 > gpurun_out/time_pool_mma.log 2>&1; echo "rc=$?" >> gpurun_out/time_pool_mma.log
timeout 600 python bench.py --steps 20 --warmup 5 --masks overlap --extras none --no-cpu > gpurun_out/bench_r02_overlap.json 2> gpurun_out/bench_r02_overlap.err; echo "rc=$?" >> gpurun_out/bench_r02_overlap.err
timeout 600 python bench.py --steps 20 --warmup 5 --masks overlap --pool-path rows --extras none --no-cpu > gpurun_out/bench_r02_overlap_rows.json 2>> gpurun_out/bench_r02_overlap.err; echo "rc=$?" >> gpurun_out/bench_r02_overlap.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pool" > gpurun_out/gpu_tests7.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests7.log
echo done

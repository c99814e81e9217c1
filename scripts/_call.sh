set -x
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err ) 2> gpurun_out/bench_r02_a.time; echo "rc=$?" >> gpurun_out/bench_r02_a.err
echo done

set -x
timeout 900 python bench.py > gpurun_out/bench37.json 2> gpurun_out/bench37.err; echo "rc=$?" >> gpurun_out/bench37.err
echo done

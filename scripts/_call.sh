set -x
timeout 900 python bench.py --extras none > gpurun_out/bench39.json 2> gpurun_out/bench39.err; echo "rc=$?" >> gpurun_out/bench39.err
echo done

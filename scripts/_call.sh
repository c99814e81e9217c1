set -x
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/bench45_n8.json 2> gpurun_out/bench45_n8.err; echo "rc=$?" >> gpurun_out/bench45_n8.err
echo done

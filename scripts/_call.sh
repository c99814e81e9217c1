set -x
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench38_n2.json 2> gpurun_out/bench38_n2.err; echo "rc=$?" >> gpurun_out/bench38_n2.err
echo done

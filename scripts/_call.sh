set -x
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err ) 2> gpurun_out/bench_r02_n8.time; echo "rc=$?" >> gpurun_out/bench_r02_n8.err
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
lscpu | head -25 > gpurun_out/lscpu_n8.txt 2>&1
echo done

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
cat gpurun_out/bench_n8.json; tail -4 gpurun_out/bench_n8.err
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -14 gpurun_out/topo.txt

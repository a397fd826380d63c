#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
cat gpurun_out/bench_n$n.json; tail -4 gpurun_out/bench_n$n.err
done

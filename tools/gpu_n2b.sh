#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err
cat gpurun_out/bench_n2_fused.json; tail -15 gpurun_out/bench_n2_fused.err

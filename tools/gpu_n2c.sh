#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/bench_c3_multi.py --docs-per-gpu 1250000 > gpurun_out/c3_n2.json 2> gpurun_out/c3_n2.err
cat gpurun_out/c3_n2.json; tail -8 gpurun_out/c3_n2.err

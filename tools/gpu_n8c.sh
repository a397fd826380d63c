#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/bench_c3_multi.py --docs-per-gpu 1250000 --queries-per-gpu 12500 > gpurun_out/c3_n8.json 2> gpurun_out/c3_n8.err
cat gpurun_out/c3_n8.json; tail -4 gpurun_out/c3_n8.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/bench_c5_multi.py --rows-per-gpu 125000 > gpurun_out/c5_n8.json 2> gpurun_out/c5_n8.err
cat gpurun_out/c5_n8.json; tail -4 gpurun_out/c5_n8.err

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/bench_c5_multi.py --rows-per-gpu 100000 > gpurun_out/c5_n2.json 2> gpurun_out/c5_n2.err
cat gpurun_out/c5_n2.json; tail -8 gpurun_out/c5_n2.err
timeout -s KILL 500 python tools/bench_c1.py > gpurun_out/c1_bulk.jsonl 2> gpurun_out/c1_bulk.err
cat gpurun_out/c1_bulk.jsonl; tail -5 gpurun_out/c1_bulk.err

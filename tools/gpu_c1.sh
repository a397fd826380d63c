#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_sha1_gpu.py tests/test_minhash_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout -s KILL 500 python tools/bench_c1.py > gpurun_out/c1_bulk.jsonl 2> gpurun_out/c1_bulk.err
cat gpurun_out/c1_bulk.jsonl; tail -5 gpurun_out/c1_bulk.err

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python tools/bench_e2e_sweep.py > gpurun_out/e2e_sweep.jsonl 2> gpurun_out/e2e_sweep.err
cat gpurun_out/e2e_sweep.jsonl; tail -5 gpurun_out/e2e_sweep.err

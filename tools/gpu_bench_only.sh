#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err

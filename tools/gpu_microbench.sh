#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 ./tools/microbench > gpurun_out/microbench.txt 2>&1
cat gpurun_out/microbench.txt

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 python tools/prof_c1.py > gpurun_out/prof_c1.txt 2>&1
cat gpurun_out/prof_c1.txt | cut -c1-150

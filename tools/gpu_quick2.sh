#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err
python -c "import json;d=json.load(open('gpurun_out/bench_auto.json'));print(d['ms_per_step'],d['value'])"; tail -3 gpurun_out/bench_auto.err
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:minhash_bulk -s 3 -c 1 -o gpurun_out/prof_twophase python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1

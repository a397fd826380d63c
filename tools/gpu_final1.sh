#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --docs 2000000 --tokens 128 --num-perm 256 > gpurun_out/bench_k256.json 2> gpurun_out/bench_k256.err
python -c "import json;d=json.load(open('gpurun_out/bench_k256.json'));print('k256 T128 2M docs',d['ms_per_step'],d['value'])"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:minhash_bulk -s 3 -c 1 -o gpurun_out/prof_twophase python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1

#!/bin/bash
# first GPU pass: parity tests, smoke, bench per kernel, microbench, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --kernel direct --no-e2e --no-cpu > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --kernel exact --no-e2e --no-cpu > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
timeout -s KILL 120 ./tools/microbench > gpurun_out/microbench.txt 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:minhash_bulk -s 3 -c 1 -o gpurun_out/prof_r1_twophase python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:minhash_bulk -s 3 -c 1 -o gpurun_out/prof_r1_direct python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --kernel direct > gpurun_out/ncu_full_direct.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_auto.json; tail -3 gpurun_out/bench_auto.err; cat gpurun_out/bench_direct.json gpurun_out/bench_exact.json; cat gpurun_out/microbench.txt

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_wmh_gpu.py tests/test_minhash_gpu.py tests/test_bbit.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/bench_configs.py --c3-docs 0 --c4-vecs 20000 --c5-rows 0 > gpurun_out/c4_a.jsonl 2> gpurun_out/c4.err; cat gpurun_out/c4_a.jsonl
timeout -s KILL 300 python tools/bench_configs.py --c3-docs 0 --c4-vecs 100000 --c5-rows 0 > gpurun_out/c4_b.jsonl 2>> gpurun_out/c4.err; cat gpurun_out/c4_b.jsonl; tail -3 gpurun_out/c4.err
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err
python -c "import json;d=json.load(open('gpurun_out/bench_auto.json'));print(d['ms_per_step'],d['value'],d['int_pipe'])"; tail -3 gpurun_out/bench_auto.err

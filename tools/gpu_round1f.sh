#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for m in 0 1 2; do
  DSK_K256_MODE=$m timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --docs 2000000 --tokens 128 --num-perm 256 > gpurun_out/bench_k256_m$m.json 2> gpurun_out/bench_k256_m$m.err
  python -c "import json;d=json.load(open('gpurun_out/bench_k256_m$m.json'));print('k256 mode',$m,d['ms_per_step'],d['value'])"; tail -2 gpurun_out/bench_k256_m$m.err
done
timeout -s KILL 300 python tools/bench_codecs.py > gpurun_out/codecs.jsonl 2> gpurun_out/codecs.err
cat gpurun_out/codecs.jsonl; tail -3 gpurun_out/codecs.err

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for v in 40 41 30 31; do
  DSK_TWO_PHASE_VARIANT=$v timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_v$v.json'));print('variant',$v,'T256',d['ms_per_step'],d['value'])"; tail -2 gpurun_out/bench_v$v.err
  DSK_TWO_PHASE_VARIANT=$v timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --docs 4000000 --tokens 64 > gpurun_out/bench_t64_v$v.json 2> gpurun_out/bench_t64_v$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_t64_v$v.json'));print('variant',$v,'T64',d['ms_per_step'],d['value'])"
done
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --docs 2000000 --tokens 128 --num-perm 256 > gpurun_out/bench_k256.json 2> gpurun_out/bench_k256.err
python -c "import json;d=json.load(open('gpurun_out/bench_k256.json'));print('k256 T128 2M docs',d['ms_per_step'],d['value'])"

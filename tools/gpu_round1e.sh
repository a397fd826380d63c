#!/bin/bash
mkdir -p gpurun_out
for occ in 4 5; do
  DSK_TWO_PHASE_OCC=$occ timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_occ$occ.json 2> gpurun_out/bench_occ$occ.err
  python -c "import json;d=json.load(open('gpurun_out/bench_occ$occ.json'));print('occ',$occ,d['ms_per_step'],d['value'])"
done
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --docs 2000000 --tokens 128 --num-perm 256 > gpurun_out/bench_k256.json 2> gpurun_out/bench_k256.err
python -c "import json;d=json.load(open('gpurun_out/bench_k256.json'));print('k256 T128 2M docs',d['ms_per_step'],d['value'])"; tail -2 gpurun_out/bench_k256.err
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --docs 4000000 --tokens 64 --num-perm 128 > gpurun_out/bench_t64.json 2> gpurun_out/bench_t64.err
python -c "import json;d=json.load(open('gpurun_out/bench_t64.json'));print('k128 T64 4M docs',d['ms_per_step'],d['value'])"; tail -2 gpurun_out/bench_t64.err
timeout -s KILL 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
cat gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.err

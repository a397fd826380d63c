#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for occ in 4 5 6; do
  DSK_TWO_PHASE_OCC=$occ timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_occ$occ.json 2> gpurun_out/bench_occ$occ.err
done
timeout -s KILL 300 python tools/bench_codecs.py > gpurun_out/codecs.jsonl 2> gpurun_out/codecs.err
tail -5 gpurun_out/pytest_gpu.log
for occ in 4 5 6; do python -c "import json;d=json.load(open('gpurun_out/bench_occ$occ.json'));print('occ',$occ,d['ms_per_step'],d['value'])"; done
cat gpurun_out/codecs.jsonl; tail -3 gpurun_out/codecs.err

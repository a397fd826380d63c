#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_minhash_gpu.py tests/test_sha1_gpu.py -x -q -m gpu 2>&1 | tail -1
for cfg in "1000000 256 128" "1000000 256 192" "1000000 256 100"; do
  set -- $cfg
  timeout -s KILL 100 python bench.py --no-cpu --no-e2e --docs $1 --tokens $2 --num-perm $3 --steps 10 2>/dev/null | python -c "
import sys,json,os
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'][12:60], round(d['ms_per_step'],3), round(d['value']/1e6,1))
"
done

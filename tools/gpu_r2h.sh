#!/bin/bash
# round 2, call H: tracking ops deferred into the next block's IMAD stream; 16-byte LSH slots
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q --timeout 1500 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
timeout -s KILL 600 python tools/bench_shapes.py > gpurun_out/r2h_shapes.jsonl 2> gpurun_out/r2h_shapes.err; echo "shapes rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2h_shapes.jsonl"):
    try:
        d = json.loads(l)
    except Exception:
        print(l[:200]); continue
    print(d["shape"], d["repeat_share"], d["ms"], round(d["frac_of_imad_floor"], 3), d["rows_identical"])
PY
timeout -s KILL 600 python tools/bench_configs.py --c3-docs 2000000 --c4-vecs 8 --c5-rows 1000 > gpurun_out/r2h_configs.jsonl 2> gpurun_out/r2h_configs.err; echo "configs rc=$?"; head -1 gpurun_out/r2h_configs.jsonl | cut -c1-500
timeout -s KILL 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(d['value'], d['ms_per_step'], d['int_pipe']['frac'], d['duplicates'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'])"

#!/bin/bash
# round 2, call D: GPU suite with the long-document path + new dedupe, shapes, ncu launch list of the bench command
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q --timeout 1500 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2d_pytest.log
timeout -s KILL 600 python tools/bench_shapes.py > gpurun_out/r2d_shapes.jsonl 2> gpurun_out/r2d_shapes.err; echo "shapes rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2d_shapes.jsonl"):
    try:
        d = json.loads(l)
    except Exception:
        print(l[:200]); continue
    print(d["shape"], d["repeat_share"], d["ms"], round(d["frac_of_imad_floor"], 3), d["rows_identical"])
PY
tail -3 gpurun_out/r2d_shapes.err

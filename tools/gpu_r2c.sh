#!/bin/bash
# round 2, call C: full GPU suite (incl. LSHBloom, Redis layout, the reference's own tests), shapes with OCC 4 / 5, bench
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q --timeout 1500 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2c_pytest.log
timeout -s KILL 600 python tools/bench_shapes.py > gpurun_out/r2c_shapes.jsonl 2> gpurun_out/r2c_shapes.err; echo "shapes rc=$?"
DSK_SIG_OCC=5 timeout -s KILL 600 python tools/bench_shapes.py > gpurun_out/r2c_shapes_occ5.jsonl 2>&1
python - <<'PY'
import json
for f in ("r2c_shapes", "r2c_shapes_occ5"):
    for l in open("gpurun_out/%s.jsonl" % f):
        try:
            d = json.loads(l)
        except Exception:
            print(l[:200]); continue
        print(f, d["shape"], d["repeat_share"], d["ms"], round(d["frac_of_imad_floor"], 3), d["rows_identical"])
PY
timeout -s KILL 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2c_bench.json

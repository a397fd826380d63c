#!/bin/bash
# round 2, call E: GPU suite, shapes (1024-token pieces), C3 LSH insert timing + ncu capture of the new insert kernel
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q --timeout 1500 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_pytest.log
timeout -s KILL 600 python tools/bench_shapes.py > gpurun_out/r2e_shapes.jsonl 2> gpurun_out/r2e_shapes.err; echo "shapes rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2e_shapes.jsonl"):
    try:
        d = json.loads(l)
    except Exception:
        print(l[:200]); continue
    print(d["shape"], d["repeat_share"], d["ms"], round(d["frac_of_imad_floor"], 3), d["rows_identical"])
PY
timeout -s KILL 600 python tools/bench_configs.py --c3-docs 2000000 --c4-vecs 4000 --c5-rows 20000 > gpurun_out/r2e_configs.jsonl 2> gpurun_out/r2e_configs.err; echo "configs rc=$?"; cut -c1-400 gpurun_out/r2e_configs.jsonl
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"lsh_insert_kernel" -c 1 -o gpurun_out/r2e_prof_lsh python tools/bench_configs.py --c3-docs 2000000 --c4-vecs 8 --c5-rows 64 > gpurun_out/r2e_ncu_lsh.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
# round 2, call G: LSH insert (CTA-tile kernel) timing + ncu capture; secondary configs
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_lsh_gpu.py tests/test_lshforest.py tests/test_lshensemble.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.log
timeout -s KILL 600 python tools/bench_configs.py --c3-docs 2000000 --c4-vecs 20000 --c5-rows 100000 > gpurun_out/r2g_configs.jsonl 2> gpurun_out/r2g_configs.err; echo "configs rc=$?"; cut -c1-700 gpurun_out/r2g_configs.jsonl; tail -3 gpurun_out/r2g_configs.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"lsh_insert_kernel" -s 1 -c 1 -o gpurun_out/r2g_prof_lsh python tools/bench_configs.py --c3-docs 2000000 --c4-vecs 8 --c5-rows 64 > gpurun_out/r2g_ncu_lsh.log 2>&1; echo "ncu rc=$?"

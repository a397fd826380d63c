#!/bin/bash
# ncu --set full captures of the secondary kernels (one launch each) + full GPU test suite
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"wmh_kernel|jaccard_topk_kernel|lsh_query_kernel|lsh_insert_kernel" -c 6 -o gpurun_out/prof_secondary python tools/bench_configs.py --c3-docs 1000000 --c4-vecs 4000 --c5-rows 20000 > gpurun_out/ncu_secondary.log 2>&1
tail -3 gpurun_out/ncu_secondary.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"lean_tile_kernel|band_keys_be_kernel|band_fingerprint_kernel" -s 8 -c 6 -o gpurun_out/prof_codecs python tools/bench_codecs.py > gpurun_out/ncu_codecs.log 2>&1
tail -3 gpurun_out/ncu_codecs.log
ls -la gpurun_out/*.ncu-rep

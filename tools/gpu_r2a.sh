#!/bin/bash
# round 2, call A: baseline bench (new CPU arm) + reference arm + ncu source-level capture of the hot kernel
mkdir -p gpurun_out
nproc > gpurun_out/r2a_nproc.txt
timeout -s KILL 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.json
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
tail -c 400 gpurun_out/r2a_bench_ref.json
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:minhash_bulk_kernel -s 3 -c 1 -o gpurun_out/r2a_prof \
  python bench.py --docs 400000 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep

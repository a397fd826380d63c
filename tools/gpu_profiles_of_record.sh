#!/bin/bash
# round 2, call K: profiles of record -- launch list of the default bench command, ncu --set full of the signature kernel at the
# bench's size, default bench line + reference arm
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; echo "ref rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print(d['value'], d['ms_per_step'], d['int_pipe']['frac'], d['duplicates']['slowdown_vs_unique'], d['e2e']['value'], d['e2e']['frac_of_copy_floor'], d['cpu_baseline']['value'], d['gpu_launches'])
d=json.load(open('gpurun_out/r2k_bench_ref.json')); print(d['value'], d['cpu_baseline']['kind'])"
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2k_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:minhash_sig_kernel -s 6 -c 1 -o gpurun_out/r2k_prof_sig \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-dups > gpurun_out/r2k_ncu_sig.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2k_*

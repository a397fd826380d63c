#!/bin/bash
# round 2, call I: kernel variants A/B on one box (same sources, -D switches), prefilter top-k, e2e repeatability
mkdir -p gpurun_out
V=datasketch_b200/variants
S="c2_aligned_1Mx256 ragged_1M_128to384 k256_2Mx128 short_ragged_4M_16to112"
: > gpurun_out/r2i_ab.txt
for rep in 1 2; do
  for v in base nopend nodefer nopend_nodefer g2; do
    DSK_B200_LIB=$V/libdsk_$v.so timeout -s KILL 300 python tools/bench_shapes.py $S 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$v', d['shape'], d['ms'], d['rows_identical'])" >> gpurun_out/r2i_ab.txt
  done
  for occ in 4 5 6; do
    DSK_SIG_OCC=$occ DSK_B200_LIB=$V/libdsk_noswp.so timeout -s KILL 300 python tools/bench_shapes.py $S 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('noswp_occ$occ', d['shape'], d['ms'], d['rows_identical'])" >> gpurun_out/r2i_ab.txt
  done
done
sort gpurun_out/r2i_ab.txt | awk '{k=$1" "$2; s[k]+=$3; n[k]++; ok[k]=ok[k] $4} END {for (k in s) print k, s[k]/n[k], ok[k]}' | sort -k2,2 -k3,3n
timeout -s KILL 900 python -m pytest tests/test_codec_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 600 python tools/bench_configs.py --c3-docs 200000 --c4-vecs 8 --c5-rows 100000 2> gpurun_out/r2i_configs.err | cut -c1-400
for i in 1 2; do timeout -s KILL 600 python bench.py --no-cpu --no-dups 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['frac_of_copy_floor'])"; done

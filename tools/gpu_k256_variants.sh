#!/bin/bash
# K > 128 code-generation variants (tools/build_variants.sh k8_*) on one box, C3's signature shape (2M x 128 tokens, K = 256) and
# a K = 192 shape; plus an ncu capture (source counters) of the u64-token kernel on documents with repeated tokens.
mkdir -p gpurun_out
: > gpurun_out/r2v_k256_variants.txt
for rep in 1; do
for lib in datasketch_b200/variants/libdsk_k8_*.so; do
  name=$(basename $lib .so)
  DSK_B200_LIB=$PWD/$lib timeout -s KILL 200 python tools/bench_shapes.py k256_2Mx128 k192_2Mx128 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print('$name', l.strip()[:200]); continue
    print('$name', d['shape'], d['ms'], d['rows_identical'])" >> gpurun_out/r2v_k256_variants.txt
done
done
cat gpurun_out/r2v_k256_variants.txt
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:minhash_sig_kernel -s 6 -c 1 -o gpurun_out/r2v_prof_sig_u64_repeats \
  python tools/bench_shapes.py u64_tokens_repeats_500kx256 > gpurun_out/r2v_ncu_u64_repeats.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -5

#!/bin/bash
# Same-box timing of every library under datasketch_b200/variants/ (built by tools/build_variants.sh from the same sources
# with different -D switches): run under gpurun, read gpurun_out/kernel_grid.txt. The signature kernel's speed moves by
# +-3% with code generation changes OUTSIDE its hot loop, so variants are only comparable on one box in one call.
mkdir -p gpurun_out
python tools/ab_libs.py datasketch_b200/variants/*.so > gpurun_out/kernel_grid.txt 2>&1
sort gpurun_out/kernel_grid.txt | awk '{k=$1; a[k]+=$5; r[k]+=$8; n[k]++; ok[k]=ok[k] $6 $9} END {for (k in a) printf "%s aligned %.4f ragged %.4f %s\n", k, a[k]/n[k], r[k]/n[k], ok[k]}' | sort -k3,3n

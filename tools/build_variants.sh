#!/bin/bash
# Kernel A/B experiments: builds the library several times from the SAME sources with different -D switches into
# datasketch_b200/variants/ (git-ignored *.so); select one at run time with DSK_B200_LIB=<path>.
#   tools/build_variants.sh "name1:-DFOO" "name2:-DFOO -DBAR" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p datasketch_b200/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --shared $flags \
    -o datasketch_b200/variants/libdsk_$name.so datasketch_b200/csrc/*.cu &
done
wait
ls -la datasketch_b200/variants/

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q --timeout 900 -k "not full_size and not large_roundtrip and not medium_batch and not random_signatures_finds and not long_documents and not vs_dict_oracle_random" > gpurun_out/sanitizer_memcheck_all.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_all.log
tail -8 gpurun_out/sanitizer_memcheck_all.log

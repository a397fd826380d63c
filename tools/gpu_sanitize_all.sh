#!/bin/bash
# compute-sanitizer over the GPU tests (the big-size tests are deselected: memcheck slows kernels ~50x)
mkdir -p gpurun_out
SKIP="not full_size and not large_roundtrip and not medium_batch and not random_signatures_finds and not vs_dict_oracle_random and not false_positive_rate and not reference_unittests and not long_documents_on_the_device and not long_documents_are_split"
timeout -s KILL 2400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q --timeout 2000 -k "$SKIP" > gpurun_out/sanitizer_memcheck_all.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_all.log
tail -6 gpurun_out/sanitizer_memcheck_all.log
timeout -s KILL 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_minhash_gpu.py -m gpu -x -q --timeout 1400 -k "ragged_golden or c1_bulk_golden or duplicates or init_matrix" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -6 gpurun_out/sanitizer_racecheck.log

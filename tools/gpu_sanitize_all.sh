#!/bin/bash
# compute-sanitizer over the GPU tests of the kernels that changed most (big-size tests deselected: memcheck slows kernels
# ~50x).  torch is paged in first: on a fresh box its first import takes longer than the sanitizer's attach timeout.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
SKIP="not full_size and not large_roundtrip and not medium_batch and not random_signatures_finds and not vs_dict_oracle_random and not false_positive_rate and not long_documents and not repeated_calls"
timeout -s KILL 1000 compute-sanitizer --launch-timeout 600 --tool memcheck --error-exitcode 7 python -m pytest \
  tests/test_minhash_gpu.py "tests/test_signature_kernel_gpu.py::test_randomised_shapes_against_the_oracle[0]" tests/test_lsh_gpu.py \
  tests/test_lsh_bloom.py tests/test_codec_gpu.py -m gpu -x -q --timeout 900 -k "$SKIP" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -6 gpurun_out/sanitizer_memcheck.log
timeout -s KILL 600 compute-sanitizer --launch-timeout 600 --tool racecheck --error-exitcode 7 python -m pytest tests/test_minhash_gpu.py -m gpu -x -q --timeout 500 -k "ragged_golden or c1_bulk_golden or duplicates or init_matrix" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -4 gpurun_out/sanitizer_racecheck.log

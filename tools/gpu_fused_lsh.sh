#!/bin/bash
# fused LSH insert (dsk_lsh_insert_tokens): GPU gate, C3 timing fused vs two-step, compute-sanitizer on the new paths
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2w_pytest_gpu.txt
timeout -s KILL 300 python tools/bench_configs.py --c3-fused --c4-vecs 0 --c5-rows 0 > gpurun_out/r2w_c3_fused.jsonl 2> gpurun_out/r2w_c3_fused.err; echo "c3 rc=$?"
cat gpurun_out/r2w_c3_fused.jsonl; tail -3 gpurun_out/r2w_c3_fused.err
timeout -s KILL 200 python tools/bench_shapes.py k256_2Mx128 c2_aligned_1Mx256 2>&1 | cut -c1-260
timeout -s KILL 240 compute-sanitizer --launch-timeout 600 --tool memcheck --error-exitcode 9 python -m pytest -x -q -m gpu "tests/test_lsh_gpu.py::test_fused_insert_from_tokens_equals_the_two_step_flow[64-params2]" "tests/test_signature_kernel_gpu.py::test_unsafe_permutations_randomised_against_the_oracle" > gpurun_out/r2w_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2w_sanitizer_memcheck.log
timeout -s KILL 300 compute-sanitizer --launch-timeout 600 --tool racecheck --error-exitcode 9 python -m pytest -x -q -m gpu "tests/test_lsh_gpu.py::test_fused_insert_from_tokens_equals_the_two_step_flow[64-params2]" > gpurun_out/r2w_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2w_sanitizer_racecheck.log

#!/bin/bash
# round 2, second session: GPU gate + numbers for the general variants of the signature kernel (u64 tokens, unsafe permutations)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2t_pytest_gpu.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python tools/bench_shapes.py c2_aligned_1Mx256 ragged_1M_128to384 u64_tokens_1Mx256 u64_tokens_1Mx256_exact_kernel \
  u64_tokens_ragged_1M_128to384 u64_tokens_repeats_500kx256 u64_tokens_long_20k_x12800 unsafe_permutation_1Mx256 \
  unsafe_permutation_1Mx256_exact_kernel repeats_500kx256 > gpurun_out/r2t_shapes.jsonl 2> gpurun_out/r2t_shapes.err; echo "shapes rc=$?"
cat gpurun_out/r2t_shapes.jsonl | cut -c1-330
timeout -s KILL 900 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err; echo "bench rc=$?"
cat gpurun_out/r2t_bench_default.json; tail -3 gpurun_out/r2t_bench_default.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:minhash_sig_kernel -s 6 -c 1 -o gpurun_out/r2t_prof_sig_u64 \
  python tools/bench_shapes.py u64_tokens_1Mx256 > gpurun_out/r2t_ncu_u64.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ | tail -12

#!/bin/bash
# round 2, call B: parity of the new signature kernel + A/B against the round-1 kernel
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest.log
timeout -s KILL 600 python tools/bench_shapes.py > gpurun_out/r2b_shapes_new.jsonl 2> gpurun_out/r2b_shapes_new.err; echo "shapes new rc=$?"
DSK_TWO_PHASE_V1=1 timeout -s KILL 600 python tools/bench_shapes.py > gpurun_out/r2b_shapes_v1.jsonl 2> gpurun_out/r2b_shapes_v1.err; echo "shapes v1 rc=$?"
DSK_SIG_OCC=5 timeout -s KILL 300 python tools/bench_shapes.py c2_aligned_1Mx256 ragged_1M_128to384 > gpurun_out/r2b_shapes_occ5.jsonl 2>&1
cat gpurun_out/r2b_shapes_new.jsonl | cut -c1-230
echo ---- v1; cat gpurun_out/r2b_shapes_v1.jsonl | cut -c1-230
echo ---- occ5; cat gpurun_out/r2b_shapes_occ5.jsonl | cut -c1-230
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:minhash_sig_kernel -s 3 -c 1 -o gpurun_out/r2b_prof \
  python bench.py --docs 400000 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"

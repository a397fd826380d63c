#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err
python -c "import json;d=json.load(open('gpurun_out/bench_auto.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['config']['host'])"; tail -3 gpurun_out/bench_auto.err
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --docs 4000000 --tokens 64 --num-perm 128 > gpurun_out/bench_t64.json 2> gpurun_out/bench_t64.err
python -c "import json;d=json.load(open('gpurun_out/bench_t64.json'));print('k128 T64 4M docs',d['ms_per_step'],d['value'])"
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --docs 2000000 --tokens 128 --num-perm 256 > gpurun_out/bench_k256.json 2> gpurun_out/bench_k256.err
python -c "import json;d=json.load(open('gpurun_out/bench_k256.json'));print('k256 T128 2M docs',d['ms_per_step'],d['value'])"
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_minhash_gpu.py -x -q -k "ragged_golden or tail_and_alignment or duplicates or u64 or init_matrix" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -6 gpurun_out/sanitizer_memcheck.log
timeout -s KILL 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_minhash_gpu.py tests/test_codec_gpu.py -x -q -k "ragged_golden and 128 or lean_pack_golden or c1_bulk" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -6 gpurun_out/sanitizer_racecheck.log

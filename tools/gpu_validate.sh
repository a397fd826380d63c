#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_pytest.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err

#!/bin/bash
# runs the reference's own unittest files (copied into the git-ignored _refcheck/ for this run only) against
# datasketch_b200 through an import shim; only the pass/fail summary is kept
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_wmh_gpu.py -x -q -m gpu 2>&1 | tail -15
cd _refcheck && PYTHONPATH=.:.. timeout -s KILL 500 python -m pytest test -v -k "not redis" -p no:cacheprovider > ../gpurun_out/refcheck.txt 2>&1
cd ..; tail -12 gpurun_out/refcheck.txt

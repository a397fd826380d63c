#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --timeout 300 $PYTEST_ARGS > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log

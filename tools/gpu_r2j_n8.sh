#!/bin/bash
# round 2, call J (8 GPUs): bench.py exactly as the driver launches it at N=8, both arms, NCCL_DEBUG=INFO
mkdir -p gpurun_out
export NCCL_DEBUG=INFO
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err; echo "bench n2 rc=$?"
wc -l gpurun_out/r2j_bench_n8.json; tail -c 2500 gpurun_out/r2j_bench_n8.json; echo
grep -c "NCCL INFO" gpurun_out/r2j_bench_n8.err; grep -m3 -E "nranks|NVLS|comm 0x" gpurun_out/r2j_bench_n8.err | cut -c1-200
tail -5 gpurun_out/r2j_bench_n8.err | cut -c1-300
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2j_bench_ref_n8.json 2> gpurun_out/r2j_bench_ref_n8.err; echo "ref n2 rc=$?"
tail -c 600 gpurun_out/r2j_bench_ref_n8.json

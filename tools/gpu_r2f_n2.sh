#!/bin/bash
# round 2, call F (2 GPUs): bench.py exactly as the driver launches it at N=2, both arms, NCCL_DEBUG=INFO
mkdir -p gpurun_out
export NCCL_DEBUG=INFO
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"
wc -l gpurun_out/r2f_bench_n2.json; tail -c 2500 gpurun_out/r2f_bench_n2.json; echo
grep -c "NCCL INFO" gpurun_out/r2f_bench_n2.err; grep -m3 -E "nranks|NVLS|comm 0x" gpurun_out/r2f_bench_n2.err | cut -c1-200
tail -5 gpurun_out/r2f_bench_n2.err | cut -c1-300
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref_n2.json 2> gpurun_out/r2f_bench_ref_n2.err; echo "ref n2 rc=$?"
tail -c 600 gpurun_out/r2f_bench_ref_n2.json

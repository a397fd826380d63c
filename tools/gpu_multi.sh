#!/bin/bash
# bench.py exactly as the driver launches it at N GPUs, both arms, NCCL_DEBUG=INFO (the log goes to stderr):
#   gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=INFO
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
wc -l gpurun_out/bench_n$N.json; tail -c 1500 gpurun_out/bench_n$N.json; echo
grep -c "NCCL INFO" gpurun_out/bench_n$N.err; grep -m3 -E "nranks|NVLS" gpurun_out/bench_n$N.err | cut -c1-200
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref n$N rc=$?"
tail -c 400 gpurun_out/bench_ref_n$N.json

mkdir -p gpurun_out
export NCCL_DEBUG=INFO
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err; echo "bench n2 rc=$?"
wc -l gpurun_out/r2u_bench_n2.json; cat gpurun_out/r2u_bench_n2.json; echo
grep -c "NCCL INFO" gpurun_out/r2u_bench_n2.err; grep -m3 -E "nranks|NVLS" gpurun_out/r2u_bench_n2.err | cut -c1-200; grep -v "NCCL INFO" gpurun_out/r2u_bench_n2.err | tail -5
unset NCCL_DEBUG
timeout -s KILL 300 python tools/bench_shapes.py u64_tokens_repeats_500kx256 repeats_500kx256 > gpurun_out/r2u_shapes_u64_repeats.jsonl 2>&1; cat gpurun_out/r2u_shapes_u64_repeats.jsonl | cut -c1-330

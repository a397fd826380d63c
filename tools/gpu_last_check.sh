mkdir -p gpurun_out
timeout -s KILL 17 python tools/gpu_last_check.py > gpurun_out/r2y_last_check.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2y_last_check.txt | tail -12

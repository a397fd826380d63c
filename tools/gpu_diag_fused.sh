mkdir -p gpurun_out
timeout -s KILL 50 python tools/diag_fused.py > gpurun_out/r2x_diag_fused.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2x_diag_fused.txt

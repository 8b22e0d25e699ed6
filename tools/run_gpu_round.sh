mkdir -p gpurun_out
T=${TAG:-r14}
timeout -s KILL 240 python tools/gpu_diag.py --kernel tc > gpurun_out/${T}_diag.txt 2>&1; echo "rc=$?" >> gpurun_out/${T}_diag.txt

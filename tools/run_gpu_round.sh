mkdir -p gpurun_out
T=${TAG:-r16}
timeout -s KILL 240 python tools/gpu_diag.py --kernel tc > gpurun_out/${T}_diag.txt 2>&1; echo "rc=$?" >> gpurun_out/${T}_diag.txt
L2HMC_LIB=$PWD/l2hmc_b200/libl2hmc_acct.so timeout -s KILL 120 python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/${T}_diag_acct.txt 2>&1

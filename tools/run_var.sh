mkdir -p gpurun_out
for rep in 1 2; do
timeout -s KILL 120 python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/late_base_$rep.txt 2>&1
L2HMC_LIB=$PWD/l2hmc_b200/libl2hmc_late.so timeout -s KILL 120 python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/late_var_$rep.txt 2>&1
done
L2HMC_LIB=$PWD/l2hmc_b200/libl2hmc_late.so timeout -s KILL 240 python tools/gpu_diag.py --kernel tc > gpurun_out/late_var_parity.txt 2>&1

mkdir -p gpurun_out
T=${TAG:-r12}
timeout -s KILL 120 python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/${T}_diag.txt 2>&1
for v in 0 20; do L2HMC_LIB=$PWD/l2hmc_b200/libl2hmc_s$v.so timeout -s KILL 120 python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/${T}_diag_s$v.txt 2>&1; done
timeout -s KILL 300 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err

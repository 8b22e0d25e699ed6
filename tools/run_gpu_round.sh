mkdir -p gpurun_out
T=${TAG:-r6}
python tools/gpu_diag.py --kernel tc > gpurun_out/${T}_diag.txt 2>&1
L2HMC_TC_GENERIC=1 python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/${T}_diag_generic.txt 2>&1
L2HMC_LIB=$PWD/l2hmc_b200/libl2hmc_acct.so python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/${T}_diag_acct.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.txt

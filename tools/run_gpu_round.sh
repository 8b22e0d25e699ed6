mkdir -p gpurun_out
T=${TAG:-r18}
timeout -s KILL 240 python tools/gpu_diag.py --kernel tc > gpurun_out/${T}_diag.txt 2>&1; echo "rc=$?" >> gpurun_out/${T}_diag.txt
L2HMC_TC_BIASG=0 timeout -s KILL 240 python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/${T}_diag_nobg.txt 2>&1
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc or golden or full_size" > gpurun_out/${T}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.txt

mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.txt
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
for nq in 2 3 4; do L2HMC_TC_NQ=$nq python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/r1_diag_nq$nq.txt 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_transition_kernel -s 3 -c 1 -f -o gpurun_out/r1_prof python bench.py --steps 2 --warmup 3 > gpurun_out/r1_ncu.log 2>&1
ls -la gpurun_out

# Round evidence: GPU tests, smoke, bench (both arms), ncu launch list and one full capture of the dominant kernel.
mkdir -p gpurun_out
T=${TAG:-final}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.txt
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout -s KILL 300 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_transition -s 3 -c 1 -f -o gpurun_out/${T}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
# phase counters: only when the accounting build was made HERE beforehand (tools/build_variants.sh acct "-DL2HMC_TC_PHASE_ACCOUNTING");
# a missing or stale L2HMC_LIB would be rebuilt on the GPU box (minutes of box time)
[ -f l2hmc_b200/libl2hmc_acct.so ] && L2HMC_LIB=$PWD/l2hmc_b200/libl2hmc_acct.so timeout -s KILL 120 python tools/gpu_diag.py --kernel tc --no-parity > gpurun_out/${T}_phase_accounting.txt 2>&1
timeout -s KILL 300 python tools/gpu_diag.py > gpurun_out/${T}_parity_timing.txt 2>&1
# training path (first-correct version): its GPU cases in the open, and one timing beside the sampling transition
timeout -s KILL 900 python -m pytest tests/test_gpu_training.py -m gpu -q > gpurun_out/${T}_pytest_training.txt 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_training.txt
timeout -s KILL 300 python tools/train_timing.py > gpurun_out/${T}_train_timing.txt 2>&1

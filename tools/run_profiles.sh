#!/bin/bash
# ncu evidence for the round (run on the GPU box through gpurun; CSV pages under gpurun_out/ -- the .ncu-rep files stay in /tmp on
# the box: gpurun brings back at most 64 MiB -- summaries copied to profiles/ here):
#   launch list of the benchmark command, and one `--set full` capture per dominant kernel with the raw page as CSV.
R=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/${R}_launches_bench.log 2>&1
for cfg in c2 c1 c3 c4; do
  ncu --set full --clock-control none --import-source on -k regex:transition -c 1 -s 1 -f -o /tmp/${R}_${cfg}_full \
      python tools/profile_target.py $cfg 3 > gpurun_out/${R}_${cfg}_ncu.log 2>&1
  ncu -i /tmp/${R}_${cfg}_full.ncu-rep --page raw --csv > gpurun_out/${R}_${cfg}_ncu_raw.csv 2>/dev/null
  ncu -i /tmp/${R}_${cfg}_full.ncu-rep --page source --csv > gpurun_out/${R}_${cfg}_ncu_source.csv 2>/dev/null
done
# config 5 (layered engine): launch list of one transition + full capture of its dominant GEMM kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${R}_c5_launches.csv \
    python tools/profile_target.py c5 1 > gpurun_out/${R}_c5_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 1 -s 40 -f -o /tmp/${R}_c5_gemm_full \
    python tools/profile_target.py c5 1 > gpurun_out/${R}_c5_ncu.log 2>&1
ncu -i /tmp/${R}_c5_gemm_full.ncu-rep --page raw --csv > gpurun_out/${R}_c5_gemm_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/${R}_*

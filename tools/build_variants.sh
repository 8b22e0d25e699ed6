#!/bin/bash
# build kernel variants side by side: tools/build_variants.sh name1 "-DX=1 -DY=2" name2 "..." ...   -> l2hmc_b200/libl2hmc_<name>.so
while [ $# -gt 1 ]; do
  name=$1; extra=$2; shift 2
  ( L2HMC_LIB=$PWD/l2hmc_b200/libl2hmc_$name.so L2HMC_NVCC_EXTRA="$extra" python -c "from l2hmc_b200 import _lib; _lib.build(force=True)" && echo "built $name" ) &
done
wait

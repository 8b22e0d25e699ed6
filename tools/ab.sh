#!/bin/bash
# A/B of kernel variants built side by side (L2HMC_LIB): headline workload only, short.  usage: tools/ab.sh lib1.so lib2.so ...
for lib in "$@"; do
  L2HMC_LIB=$PWD/l2hmc_b200/$lib python bench.py --no-other-configs --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
p=l['parity']
print('$lib', 'ms=%.4f value=%.4g e2e=%.4g kernel=%s | Lx=%.2e Lv=%.2e px=%.2e pxmean=%.2e' % (l['ms_per_step'], l['value'], l['e2e']['value'], l['config']['kernel'], p['Lx'], p['Lv'], p['px_max'], p['px_mean']))"
done

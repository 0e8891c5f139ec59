#!/bin/bash
# whole-program leg with the seam's phase stamps, at the default 20 M sites and at config 1's 50 M
set -u
OUT=gpurun_out
export BSGPU_SEAM_TIMING=1
for sz in 2e7 5e7; do
  SECONDS=0
  timeout 900 python bench.py --sites 1e8 --steps 2 --warmup 3 --no-cpu --legs binary --binary-sites $sz > $OUT/r02e_binary_$sz.json 2> $OUT/r02e_binary_$sz.err
  echo "binary leg $sz exit $? after $SECONDS s"
  python -c "
import json; d=json.load(open('$OUT/r02e_binary_$sz.json'))['full_binary']
if 'error' in d: print(d)
else:
  for k in ('cpu_reference_binary','gpu_seam_C','gpu_seam_D'):
    v=d[k]; print(k, round(v['wall_s'],2),'s wall', round(v['user_s'],2), round(v['sys_s'],2), '%.3g sites/s'%v['value']); [print('   ',x) for x in v.get('seam_timing',[])]
  print(d['parity'], d['speedup_wall'])"
done

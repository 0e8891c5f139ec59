#!/bin/bash
# bulk input of the wide seams: parity tests, then the whole-program leg with phase stamps (bulk against the record loop)
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_binary.py -x -q -s -k "several or seed3 or 3-extra0" 2>&1 | grep -v "^Processing chromosome\|^Warning not found\|^Loading\|^ctg loaded\|^len =\|^Read in\|dbSNP index\|^Opened\|^Additional\|^Sequence index\|^Finished\|^bs_call fin\|^Writing out" | tail -45
export BSGPU_SEAM_TIMING=1
for mode in bulk loop; do
  if [ $mode = loop ]; then export BSGPU_SEAM_BULK=0; else unset BSGPU_SEAM_BULK; fi
  for sz in 2e7 5e7; do
    timeout 900 python bench.py --sites 1e8 --steps 2 --warmup 3 --no-cpu --legs binary --binary-sites $sz > $OUT/r02e_binary_${mode}_$sz.json 2> $OUT/r02e_binary_${mode}_$sz.err
    echo "binary leg $mode $sz exit $?"
    python -c "
import json; d=json.load(open('$OUT/r02e_binary_${mode}_$sz.json'))['full_binary']
if 'error' in d: print(d)
else:
  for k in ('cpu_reference_binary','gpu_seam_C','gpu_seam_D'):
    v=d[k]; print(k, round(v['wall_s'],2),'s wall', round(v['user_s'],2), round(v['sys_s'],2), '%.3g sites/s'%v['value']); [print('   ',x) for x in v.get('seam_timing',[])]
  print(d['parity'], d['speedup_wall'])"
  done
done

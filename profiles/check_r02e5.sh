#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_session.py tests/test_gpu_seams.py -x -q 2>&1 | tail -3
export BSGPU_SEAM_TIMING=1
for mode in bulk loop; do
  if [ $mode = loop ]; then export BSGPU_SEAM_BULK=0; else unset BSGPU_SEAM_BULK; fi
  for sz in 2e7; do
    timeout 900 python bench.py --sites 1e8 --steps 2 --warmup 3 --no-cpu --legs binary --binary-sites $sz > $OUT/r02e_binary2_${mode}_$sz.json 2> $OUT/r02e_binary2_${mode}_$sz.err
    echo "binary leg $mode $sz exit $?"
    python -c "
import json; d=json.load(open('$OUT/r02e_binary2_${mode}_$sz.json'))['full_binary']
if 'error' in d: print(d)
else:
  for k in ('cpu_reference_binary','gpu_seam_C','gpu_seam_D'):
    v=d[k]; print(k, round(v['wall_s'],2),'s wall', v['wall_s_all_runs'], '%.3g sites/s'%v['value'], ' | '.join(v.get('seam_timing',[])))
  print(d['speedup_wall'])"
  done
done
unset BSGPU_SEAM_BULK
timeout 900 python bench.py --sites 1e8 --steps 2 --warmup 3 --no-cpu --legs binary --binary-sites 5e7 > $OUT/r02e_binary2_bulk_5e7.json 2> $OUT/r02e_binary2_bulk_5e7.err
python -c "
import json; d=json.load(open('$OUT/r02e_binary2_bulk_5e7.json'))['full_binary']
for k in ('cpu_reference_binary','gpu_seam_C','gpu_seam_D'):
    v=d[k]; print('5e7', k, round(v['wall_s'],2),'s wall', v['wall_s_all_runs'], '%.3g sites/s'%v['value'], ' | '.join(v.get('seam_timing',[])))
print(d['speedup_wall'], d['parity'])"

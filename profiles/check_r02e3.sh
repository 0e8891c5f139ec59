#!/bin/bash
# block path: serial (one stream) against overlapped (gather of part k + 1 next to the model of part k)
set -u
OUT=gpurun_out
timeout 600 python bench.py --sites 1e8 --steps 5 --warmup 3 --no-cpu --legs block > $OUT/r02e_block.json 2> $OUT/r02e_block.err
echo "exit $?"; tail -3 $OUT/r02e_block.err
python -c "
import json; b=json.load(open('$OUT/r02e_block.json'))['block_path']
if 'error' in b: print(b)
else:
  for k in ('default','pileup_only','fused_variant','overlapped_variant','serial_variant'):
    if k in b: print(k, round(b[k]['ms'],3),'ms', '%.4g'%b[k]['sites_per_s'], round(b[k]['roofline']['frac'],4), b[k].get('gt_vcf_bytes_identical_to_default'))
  d=b['deep_panel']; print('deep', d['pileup_only']['ms'], d['default']['ms'])"

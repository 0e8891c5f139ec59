#!/bin/bash
# Round 2, third session: whole-program binaries on the GPU, transport-specialised k_call_sites A/B, the new bench leg.
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_binary.py -x -q -s 2>&1 | grep -v "^Processing chromosome\|^Warning not found" | tail -40
for v in generic special; do
  if [ $v = generic ]; then export BSGPU_CALL_GENERIC=1; else unset BSGPU_CALL_GENERIC; fi
  timeout 300 python bench.py --sites 2.5e8 --steps 4 --warmup 3 --no-cpu --legs none > $OUT/ab_$v.json 2> $OUT/ab_$v.err
  python -c "
import json; d=json.load(open('$OUT/ab_$v.json')); print('$v', d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'])"
done
unset BSGPU_CALL_GENERIC
BSGPU_BULK=10 timeout 300 python bench.py --sites 2.5e8 --steps 4 --warmup 3 --no-cpu --legs none > $OUT/ab_async.json 2> $OUT/ab_async.err
python -c "
import json; d=json.load(open('$OUT/ab_async.json')); print('special cp.async', d['value'], d['roofline']['frac'])"
BSGPU_BULK=10 BSGPU_CALL_GENERIC=1 timeout 300 python bench.py --sites 2.5e8 --steps 4 --warmup 3 --no-cpu --legs none > $OUT/ab_async_g.json 2> $OUT/ab_async_g.err
python -c "
import json; d=json.load(open('$OUT/ab_async_g.json')); print('generic cp.async', d['value'], d['roofline']['frac'])"
SECONDS=0
timeout 600 python bench.py --sites 1e8 --steps 2 --warmup 3 --no-cpu --legs binary > $OUT/r02e_binary.json 2> $OUT/r02e_binary.err
echo "binary leg exit $? after $SECONDS s"; tail -4 $OUT/r02e_binary.err
python -c "
import json; d=json.load(open('$OUT/r02e_binary.json')); print(json.dumps(d['full_binary'], indent=1)[:3500])"

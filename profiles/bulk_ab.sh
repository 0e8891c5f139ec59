# A/B of the likelihood kernel's input prefetch: BSGPU_BULK bits (1 bulk load, 2 bulk store, 8 cp.async prefetch)
set -u
for b in 3 10 3 10; do
  BSGPU_BULK=$b python bench.py --legs block --steps 5 --warmup 3 --no-cpu > gpurun_out/bulk_$b.json 2> gpurun_out/bulk_$b.err
  python - <<PY
import json
for line in open('gpurun_out/bulk_$b.json'):
    if line.startswith('{'):
        d=json.loads(line); print('BULK', $b, 'value %.5g'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'block', (d.get('block_path') or {}).get('default_two_kernels'))
PY
done
BSGPU_BULK=10 BSGPU_DECODE_STREAM=1 timeout 200 python bench.py --legs genome --genome-sessions 2 --genome-scale 16 --sites 1e8 --steps 2 --warmup 1 --no-cpu > gpurun_out/bulk10_2l.json 2> gpurun_out/bulk10_2l.err
grep -h "genome: timed\|illegal" gpurun_out/bulk10_2l.err

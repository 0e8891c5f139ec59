#!/bin/bash
# three-phase certain-start scan on the device: parity (BSGPU_CHECK_SCAN inside the test; whole reader module with the device's mask), then the BAM leg both ways
set -u
timeout 200 python -m pytest tests/test_gpu_reader.py -x -q -k "device_scan" 2>&1 | tail -2
BSGPU_DEVICE_SCAN=1 timeout 300 python -m pytest tests/test_gpu_reader.py tests/test_gpu_session.py -x -q -k "not decode_records_golden" 2>&1 | tail -2
for v in host device; do
  if [ $v = device ]; then export BSGPU_DEVICE_SCAN=1; else unset BSGPU_DEVICE_SCAN; fi
  timeout 200 python bench.py --sites 1e8 --steps 2 --warmup 3 --no-cpu --legs bam --bam-sites 5e7 > gpurun_out/r02g_bam_$v.json 2> gpurun_out/r02g_bam_$v.err
  python -c "
import json; b=json.load(open('gpurun_out/r02g_bam_$v.json'))['bam_path']
print('$v', b.get('error') or ('%.4g %.4g' % (b['e2e']['value'], b['to_bcf_records']['value']), b['stage_s_per_step'], b['e2e']['d2h_bytes_per_step']))"
done

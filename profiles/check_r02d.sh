set -u
timeout 900 python -m pytest tests/test_gpu_writer.py tests/test_gpu_parity.py tests/test_gpu_site_stats.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --legs e2e,writer,genome --genome-scale 16 --sites 1e8 --steps 2 --warmup 3 > gpurun_out/r02d_q.json 2> gpurun_out/r02d_q.err
echo "bench exit $?"; tail -5 gpurun_out/r02d_q.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_q.json'))
print('e2e', d['e2e'])
w=d['writer_path']; print('writer', w['resident'], w['e2e']['value'], w.get('cpu_baseline',{}).get('parity_records_checked'))
g=d['genome_path']; print('genome', g['value'], g['pass_s'], g.get('parity'), g.get('dbsnp_entries_rank0'), g['bcf_records'])
PY

set -u
timeout 900 python -m pytest tests/test_gpu_writer.py tests/test_gpu_site_stats.py tests/test_gpu_session.py -x -q 2>&1 | tail -4
for w in fused split; do
  BSGPU_WRITER=$w timeout 600 python bench.py --legs writer --sites 1e8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02d_w_$w.json 2> gpurun_out/r02d_w_$w.err
  echo "$w exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02d_w_$w.json')); w=d['writer_path']; print(w['resident'], w['e2e']['value'])"
done

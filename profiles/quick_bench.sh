python bench.py --sites 6.4e7 --steps 2 --warmup 3 --no-cpu --e2e-sites 1e6 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); b=d['block_path']
print('call', d['value'], d['roofline']['frac'])
for k in ('default','pileup_only','fused_variant'): print(k, b[k]['sites_per_s'], b[k]['roofline']['frac'])
"; tail -3 gpurun_out/bench_q.err

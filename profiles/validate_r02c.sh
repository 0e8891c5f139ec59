set -u
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_gputests.log
tail -5 gpurun_out/r02c_gputests.log
for k in 1 2 3; do
  timeout 200 python bench.py --legs genome,bam --bam-sites 8e6 --genome-scale 16 --sites 1e8 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02c_g$k.json 2> gpurun_out/r02c_g$k.err
  grep -h "genome: timed\|illegal\|bam: " gpurun_out/r02c_g$k.err
done

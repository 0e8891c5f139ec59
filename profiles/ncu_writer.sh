#!/bin/bash
# one ncu --set full capture of each writer kernel (after the plain command has run once), plus the launch list
OUT=${1:-prof_writer}
SMALL="--sites 1.6e7 --steps 1 --no-cpu --e2e-sites 4e6 --fused-sites 4e6 --bam-sites 1e6"
python bench.py $SMALL > gpurun_out/plain_w.log 2>&1 || { tail -5 gpurun_out/plain_w.log; exit 1; }
for k in k_bcf_measure k_bcf_emit k_bcf_calls; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/${OUT}_$k -f python bench.py $SMALL > gpurun_out/ncu_w.log 2>&1
  tail -1 gpurun_out/ncu_w.log
done

#!/bin/bash
# one ncu --set full capture of the likelihood kernel (after the plain command has run once)
SMALL="--sites 3.2e7 --steps 1 --no-cpu --e2e-sites 1e6 --fused-sites 8e6 --bam-sites 1e6"
OUT=${1:-prof_call}
python bench.py $SMALL > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_call_sites -s 20 -c 1 -o gpurun_out/$OUT -f python bench.py $SMALL > gpurun_out/ncu_call.log 2>&1
tail -2 gpurun_out/ncu_call.log

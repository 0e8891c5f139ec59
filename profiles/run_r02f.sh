#!/bin/bash
# Round 2, final code of the third session (kernels as in r02e; host side: bulk input, session fix): full GPU suite, default bench,
# reference arm, launch list.  ncu only after the plain runs exit 0.
set -u
OUT=gpurun_out
TAG=r02f
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gputests.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_gputests.log
tail -3 $OUT/${TAG}_gputests.log
SECONDS=0
python bench.py > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench_full.err; exit 1; }
echo "default bench took $SECONDS s"
SECONDS=0
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
echo "reference arm took $SECONDS s"
SMALL="--legs e2e,block,bam,writer,genome --sites 3.2e7 --steps 1 --no-cpu --e2e-sites 1e6 --fused-sites 8e6 --bam-sites 2e6 --deep-sites 2.5e6 --genome-scale 256"
python bench.py $SMALL > $OUT/plain.log 2>&1 || { echo "small bench failed"; tail -5 $OUT/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $SMALL > $OUT/ncu_launches.log 2>&1
ls -la $OUT | tail -6

#!/bin/bash
# Round 2.  Run on the GPU box (gpurun -- bash profiles/run_profile_r02.sh [tag]): plain runs first, ncu only after they exit 0.
set -u
OUT=gpurun_out
TAG=${1:-r02}
mkdir -p $OUT
SMALL="--sites 3.2e7 --steps 1 --no-cpu --e2e-sites 1e6 --fused-sites 8e6 --bam-sites 2e6 --deep-sites 2.5e6 --genome-scale 256"
python bench.py > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench_full.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python bench.py $SMALL > $OUT/plain.log 2>&1 || { echo "small bench failed"; tail -5 $OUT/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $SMALL > $OUT/ncu_launches.log 2>&1
# DRAM traffic of the headline kernel AT BENCH SIZE (62.5 M sites per launch): two counters, one pass, no replay
BIG="--sites 1e9 --steps 1 --warmup 3 --no-cpu --legs none"
python bench.py $BIG > $OUT/plain_big.log 2>&1 || { echo "big bench failed"; tail -5 $OUT/plain_big.log; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_call_sites -s 50 -c 4 --csv --log-file $OUT/${TAG}_traffic_call_bench.csv python bench.py $BIG > $OUT/ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_call_sites -s 20 -c 2 -o $OUT/${TAG}_prof_call -f python bench.py $SMALL > $OUT/ncu_call.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pileup_tile -s 4 -c 2 -o $OUT/${TAG}_prof_pile -f python bench.py $SMALL > $OUT/ncu_pile.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_decode_records|k_name_ids|k_normalise" -c 6 -o $OUT/${TAG}_prof_reader -f python bench.py $SMALL > $OUT/ncu_reader.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bcf_emit|k_bcf_measure|k_bcf_calls" -s 9 -c 3 -o $OUT/${TAG}_prof_writer -f python bench.py $SMALL > $OUT/ncu_writer.log 2>&1
ls -la $OUT | tail -14

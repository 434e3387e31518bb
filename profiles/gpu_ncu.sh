#!/bin/bash
# ncu-only session: launch list of the bench command and one --set full capture of the hot kernel.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:yaman4_sweep -s 3 -c 1 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
cat $OUT/plain2_$TAG.log | tail -1 | cut -c1-600
timeout 300 tools/tune_yaman4.bin 37.0 > $OUT/tune_$TAG.txt 2>&1; cat $OUT/tune_$TAG.txt
ls -la $OUT

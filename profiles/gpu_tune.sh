#!/bin/bash
# Tuning session: operand-bandwidth probe, launch-shape sweep, then GPU tests and the bench line.
set -u
TAG=${1:-r1b}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 tools/dfma_probe.bin > $OUT/dfma_probe_$TAG.txt 2>&1; echo "probe rc=$?"; cat $OUT/dfma_probe_$TAG.txt
timeout 600 tools/tune_yaman4.bin 37.0 > $OUT/tune_$TAG.txt 2>&1; echo "tune rc=$?"; cat $OUT/tune_$TAG.txt
python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest_gpu_$TAG.log
python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err

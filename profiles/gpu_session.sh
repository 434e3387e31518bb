#!/bin/bash
# One gpurun session: GPU parity tests, smoke, the bench line, then the ncu launch list and one
# --set full capture of the dominant kernel (each only after the same command exited 0 without ncu).
# Usage (from the repo root, on the GPU box):  bash profiles/gpu_session.sh <tag> [skip-tests]
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
if [ "${2:-}" != "skip-tests" ]; then
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log
tail -2 $OUT/smoke_$TAG.log
fi
python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:yaman4_sweep -s 3 -c 1 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_summary.py $OUT/prof_$TAG.ncu-rep $OUT/ncu_yaman4_sweep_$TAG.csv > /dev/null 2>&1; echo "summary rc=$?"
ls -la $OUT | tail -15

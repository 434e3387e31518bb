#!/bin/bash
# ncu --set full captures of the secondary kernels (trace write-out, N-wave comb and table kernels).
set -u
OUT=gpurun_out
mkdir -p $OUT
for c in trace comb table; do
  case $c in trace) K=yaman4_fast;; comb) K=nwave_comb;; table) K=nwave_rk4;; esac
  python tools/profile_cases.py $c > $OUT/plain_$c.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/prof_$c python tools/profile_cases.py $c > $OUT/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
ls -la $OUT | grep prof_

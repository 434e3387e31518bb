#!/bin/bash
# tools/fact_thread_sweep.sh -- threads per point of the factored-table N-wave kernel (FPA_FACT_THREADS) against
# plan size and batch size; the last line of every group is the library's own choice.
# usage (GPU box): bash tools/fact_thread_sweep.sh > profiles/r2_fact_thread_sweep.txt
for N in 8 21 64 128; do
  for T in 64 128 256 512 auto; do
    if [ $T = auto ]; then unset FPA_FACT_THREADS; else export FPA_FACT_THREADS=$T; fi
    echo "N=$N T=$T $(TABLE_SKIP_PLAIN=1 TABLE_N=$N python tools/table_bench.py 1 148 1184 4736 2>&1 | grep '^B=' | sed 's/| comb.*//; s/ pt.steps\/s//; s/| factored//' | tr '\n' ' ')"
  done
done

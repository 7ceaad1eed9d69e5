#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" REPS=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:umma_ozaki --csv --log-file gpurun_out/dbg_$name.csv python profiles/gemm_one.py > gpurun_out/dbg_$name.log 2>&1; echo "$name: $(grep -E 'umma_ozaki' gpurun_out/dbg_$name.csv | awk -F'","' '{print $NF}' | tail -2 | tr '\n' ' ')"; }
for k in 4096; do for d in 7 11 19 31; do run pair_K${k}_dbg$d RDB_OZAKI_PAIR=1 RDB_OZAKI_DBG=$d K=$k; done; done

#!/bin/bash
# kernel-only duration (ncu) of the int8-slice GEMM variants at 4096^3
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" REPS=2 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg --clock-control none -k regex:umma_ozaki --csv --log-file gpurun_out/var_$name.csv python profiles/gemm_one.py > gpurun_out/var_$name.log 2>&1; echo "$name: $(grep -E 'umma_ozaki' gpurun_out/var_$name.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -4 | tr '\n' ' ')"; }
run pair RDB_OZAKI_PAIR=1
run cl2 RDB_OZAKI_PAIR=0 RDB_OZAKI_CLUSTER=2
run cl4 RDB_OZAKI_PAIR=0 RDB_OZAKI_CLUSTER=4
run cl1_kb64 RDB_OZAKI_PAIR=0 RDB_OZAKI_CLUSTER=1
run cl1_kb32 RDB_OZAKI_PAIR=0 RDB_OZAKI_CLUSTER=1 RDB_OZAKI_KB=32

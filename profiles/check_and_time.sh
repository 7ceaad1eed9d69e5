#!/bin/bash
# correctness of the int8-slice GEMM variants (pair_check.py) and kernel-only time at 4096^3 (ncu)
mkdir -p gpurun_out
for v in "RDB_OZAKI_PAIR=1"; do
  echo "== $v"; env $v timeout 300 python profiles/pair_check.py 2>&1 | tail -3
  env $v REPS=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:umma_ozaki --csv --log-file gpurun_out/ct.csv python profiles/gemm_one.py > gpurun_out/ct.log 2>&1
  grep -E 'umma_ozaki' gpurun_out/ct.csv | awk -F'","' '{print $5, $NF}' | tail -2
done

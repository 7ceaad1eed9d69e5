#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
  RDB_OZAKI_PAIR=$v REPS=1 ncu --set full --clock-control none --import-source on -k regex:umma_ozaki -s 2 -c 1 -f -o gpurun_out/gemm_pair$v python profiles/gemm_one.py > gpurun_out/ncu_pair$v.log 2>&1
  tail -2 gpurun_out/ncu_pair$v.log
done

#!/bin/bash
# kernel-only durations of the int8-slice GEMM variants over a K sweep (ncu launch list), pair (cta_group::2) vs cluster multicast
mkdir -p gpurun_out
for v in 1 0; do
  RDB_OZAKI_PAIR=$v ncu --metrics gpu__time_duration.sum --clock-control none -k regex:umma_ozaki --csv --log-file gpurun_out/ksweep_pair$v.csv python profiles/gemm_k_sweep.py > gpurun_out/ksweep_pair$v.log 2>&1
done
python - <<'PY'
import csv, collections
for v in (1, 0):
    rows = [r for r in csv.reader(open(f"gpurun_out/ksweep_pair{v}.csv")) if len(r) > 5 and r[0].isdigit()]
    print("PAIR", v, [(r[4][:24], r[-1]) for r in rows][::5])
PY

#!/bin/bash
# Round-2 ncu evidence after plane trimming (run under gpurun, one GPU): (1) the launch list of the bench command, (2) full captures of
# four consecutive int8-slice GEMM launches of the headline config (dense forward and trimmed adjoint launches), (3) the slicing passes.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extra > gpurun_out/cap_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2b.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extra > gpurun_out/cap_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_ozaki2c -s 12 -c 6 -f -o gpurun_out/prof_ozaki2c_r2b \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extra > gpurun_out/cap_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_slice_i8|k_row_absmax" -s 16 -c 6 -f -o gpurun_out/prof_slice_r2b \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extra > gpurun_out/cap_slice.log 2>&1
ls -la gpurun_out/*.ncu-rep

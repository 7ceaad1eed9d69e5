#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): (1) the launch list of the bench command, (2) a full capture of the dominant
# kernel of the headline config, (3) the two kernels of the fused hessian!, (4) the fused elementwise run of the unfused cfg 4 spelling.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/cap_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extra > gpurun_out/cap_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_ozaki2c -s 6 -c 1 -f -o gpurun_out/prof_ozaki2c_r2 \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-extra > gpurun_out/cap_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fill_block_pdl|k_hessian_patch" -s 20 -c 2 -f -o gpurun_out/prof_hessian_r2 \
    python profiles/hess_probe.py > gpurun_out/cap_hess.log 2>&1
ls -la gpurun_out/*.ncu-rep

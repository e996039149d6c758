#!/bin/bash
# Round-2 evidence run (one GPU): bench line, ncu launch list of the same command, one
# `--set full` capture of a bench-step launch of the top kernel.  Run under gpurun from the
# repo root; everything lands in gpurun_out/.
set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_1gpu.json 2> gpurun_out/r02_bench_ref_1gpu.err
python bench.py --config crambin --replicas 1024 --steps 3 --warmup 3 --no-campaign > gpurun_out/r02_bench_crambin_1gpu.json 2> gpurun_out/r02_bench_crambin_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-campaign --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hmc_ensemble --launch-skip 2 -c 1 -f \
    -o gpurun_out/r02_bench_step python bench.py --steps 1 --warmup 1 --no-campaign --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_bench_step.ncu-rep

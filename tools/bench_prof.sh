#!/bin/bash
# ncu evidence for the main bench (run after `python bench.py` has exited 0 without ncu): launch list of the whole command,
# then one --set full capture of the dominant kernel
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bp_plain.json 2> gpurun_out/bp_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-strong > gpurun_out/bp_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mala_fast_kernel -s 2 -c 1 -o gpurun_out/prof_mala_r02 -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-strong > gpurun_out/bp_ncu2.log 2>&1
ls -la gpurun_out/prof_mala_r02.ncu-rep gpurun_out/launches_r02.csv

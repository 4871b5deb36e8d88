#!/bin/bash
# ncu captures of the tensor-core kernels (run AFTER the same commands have exited 0 without ncu): flow pass, fused jump
mkdir -p gpurun_out
python tools/bench_flow.py --dim 100 --dtype bf16 --layers 4 --hidden 256 > gpurun_out/tcp_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:flow_tc_kernel -s 4 -c 1 -o gpurun_out/prof_tc_r02c -f \
    python tools/bench_flow.py --dim 100 --dtype bf16 --layers 4 --hidden 256 > gpurun_out/tcp_ncu1.log 2>&1
python tools/bench_configs.py --only "wide-flow imh" > gpurun_out/tcp_plain2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:jump_tc_kernel -s 4 -c 1 -o gpurun_out/prof_jumptc_r02c -f \
    python tools/bench_configs.py --only "wide-flow imh" > gpurun_out/tcp_ncu2.log 2>&1
ls -la gpurun_out/prof_tc_r02c.ncu-rep gpurun_out/prof_jumptc_r02c.ncu-rep
cat gpurun_out/tcp_plain.log; cut -c1-300 gpurun_out/tcp_plain2.log

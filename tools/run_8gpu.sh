#!/bin/bash
# 8-GPU records of BASELINE configs C4 / C5 at their stated sizes (2^20 chains over the GPUs of one box) and the main bench.
# usage: gpurun --gpus 8 -- tools/run_8gpu.sh [G]
G=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out/configs_r02_${G}gpu.jsonl
: > $O
$T tools/run_c5.py --strategy imh --potential rb --dim 100 --iters 20 >> $O 2> gpurun_out/run8_c4.err
$T tools/run_c5.py --strategy adaptive_imh --potential rb --dim 100 --iters 20 >> $O 2> gpurun_out/run8_c4a.err
$T tools/run_c5.py --no-fit --iters 3 >> $O 2> gpurun_out/run8_c5n.err
$T tools/run_c5.py --iters 3 >> $O 2> gpurun_out/run8_c5.err
$T tools/run_c5.py --strategy neutra_hmc --potential fn --dim 100 --iters 3 >> $O 2> gpurun_out/run8_c3.err
$T bench.py --gpus $G --steps 20 --warmup 3 > gpurun_out/bench_r02c_${G}gpu.json 2> gpurun_out/bench_r02c_${G}gpu.err
grep -v "^NCCL" $O > $O.tmp; mv $O.tmp $O
cat $O gpurun_out/bench_r02c_${G}gpu.json | cut -c1-700
tail -n 5 gpurun_out/run8_*.err

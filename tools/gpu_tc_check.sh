#!/bin/bash
# GPU check of the pipelined tensor-core flow kernel: parity tests, then throughput at the verdict's shape, then a timeline
mkdir -p gpurun_out
rm -f gpurun_out/tc_bench.jsonl gpurun_out/tc_bench.err
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -x -q > gpurun_out/tc_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tc_tests.log
tail -12 gpurun_out/tc_tests.log
for args in "--layers 4 --hidden 256" "--layers 4 --hidden 256 --op inverse" "--layers 4 --hidden 256 --op log_prob" "--layers 2 --hidden 64" "--layers 4 --hidden 128" "--layers 2 --hidden 16" $EXTRA_BENCH; do
  timeout 120 python tools/bench_flow.py --dim 100 --dtype bf16 $args >> gpurun_out/tc_bench.jsonl 2>> gpurun_out/tc_bench.err
done
cat gpurun_out/tc_bench.jsonl
tail -5 gpurun_out/tc_bench.err
if [ -f nfmc_b200/libnfmc_b200_trace.so ]; then
  timeout 120 python tools/tc_trace.py 256 4 > gpurun_out/trace_256.txt 2>&1
fi

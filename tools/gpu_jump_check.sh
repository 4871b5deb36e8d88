#!/bin/bash
# GPU check of the fused tensor-core jump: parity tests, then wide jump_mala / imh throughput through the public API
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q > gpurun_out/tc_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tc_tests.log
tail -15 gpurun_out/tc_tests.log
timeout 300 python tools/bench_configs.py --only wide > gpurun_out/configs_wide.jsonl 2> gpurun_out/configs_wide.err
cat gpurun_out/configs_wide.jsonl; tail -3 gpurun_out/configs_wide.err

#!/bin/bash
# quick throughput check of the tensor-core paths: flow pass, fused IMH iteration, NeuTra-HMC
for args in "--layers 4 --hidden 256" "--layers 4 --hidden 256 --op inverse" "--layers 4 --hidden 256 --op log_prob" "--layers 2 --hidden 64"; do
  timeout 120 python tools/bench_flow.py --dim 100 --dtype bf16 $args 2>&1 | cut -c1-200
done
timeout 300 python tools/bench_configs.py --only "wide-flow imh" 2>&1 | python -c "import sys,json; [print(j['config'][:60], '%.4g' % j['chain_steps_per_s']) for j in map(json.loads, sys.stdin)]"
timeout 300 python tools/bench_configs.py --only "C3-wide" 2>&1 | python -c "import sys,json; [print(j['config'][:60], '%.4g' % j['chain_steps_per_s']) for j in map(json.loads, sys.stdin)]"

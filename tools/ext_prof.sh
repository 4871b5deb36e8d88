#!/bin/bash
# ncu launch list of the reference README's example with a lambda target (external-target path): which kernels run, and
# how the step splits between the nfmc_ext_* kernels, the flow kernels and autograd's kernels for the callable
mkdir -p gpurun_out
cat > /tmp/ext_demo.py <<'PY'
import torch, nfmc_b200
torch.manual_seed(0)
out = nfmc_b200.sample(lambda x: torch.sum(x ** 2, dim=1), event_shape=(25,), strategy="jump_mala", n_chains=4096, n_iterations=3,
                       show_progress=False, inner_param_kwargs=dict(n_iterations=10))
print(out.samples.shape, float(out.statistics.acceptance_rate))
PY
PYTHONPATH=$PWD python /tmp/ext_demo.py > gpurun_out/ext_demo.log 2>&1 || exit 1
PYTHONPATH=$PWD ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_external.csv \
    python /tmp/ext_demo.py > gpurun_out/ext_demo_ncu.log 2>&1
tail -2 gpurun_out/ext_demo.log

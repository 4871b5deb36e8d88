#!/bin/bash
# ncu --set full of the default-flow jump kernels (2^20 chains, d = 100): jump_propose_accept_kernel (IMH iteration) and flow_pass_kernel
mkdir -p gpurun_out
export PYTHONPATH=$PWD
cat > /tmp/imh_one.py <<'PY'
import torch, nfmc_b200
from nfmc_b200.potentials import make_potential
from nfmc_b200.flow import create_flow_object
torch.manual_seed(0)
d = 100
f = create_flow_object("realnvp", (d,))
with torch.no_grad():
    for p in f.parameters():
        p.add_(0.05 * torch.randn_like(p))
s = nfmc_b200.create_sampler(make_potential("g0", (d,)), flow=f, strategy="jump_mala", param_kwargs=dict(n_iterations=3, store_samples=False), inner_param_kwargs=dict(n_iterations=1))
x0 = torch.randn(1 << 20, d, device="cuda") * 0.5
for _ in range(2):
    out = s.sample(x0, show_progress=False)
torch.cuda.synchronize()
print(out.statistics.acceptance_rate)
PY
python /tmp/imh_one.py > gpurun_out/imh_one.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:jump_propose_accept -s 3 -c 1 -o gpurun_out/prof_jpa_r02 -f python /tmp/imh_one.py > gpurun_out/jpa_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flow_pass_kernel -s 3 -c 1 -o gpurun_out/prof_fpass_r02 -f python /tmp/imh_one.py > gpurun_out/fpass_ncu.log 2>&1
ls -la gpurun_out/prof_jpa_r02.ncu-rep gpurun_out/prof_fpass_r02.ncu-rep

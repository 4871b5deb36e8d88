#!/bin/bash
# ncu --set full of one row-tile fp32 flow pass (flow_train_wide_kernel, PASS mode, weights packed transposed) at the reference's
# deep conditioner shape: d = 100, Lc = 10, n_layers = 5, n_hidden = 100, 65 536 rows
mkdir -p gpurun_out
export PYTHONPATH=$PWD
cat > /tmp/wide_one.py <<'PY'
import torch
from nfmc_b200.flow import create_flow_object
f = create_flow_object('realnvp%{"n_layers": 10, "conditioner_kwargs": {"n_layers": 5, "n_hidden": 100}}', (100,)).to("cuda")
x = torch.randn(1 << 16, 100, device="cuda")
for _ in range(3):
    z, ld = f.bijection.forward(x)
torch.cuda.synchronize()
print(float(ld.mean()))
PY
python /tmp/wide_one.py > gpurun_out/wide_one.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:flow_train_wide -s 2 -c 1 -o gpurun_out/prof_widepass_r02 -f python /tmp/wide_one.py > gpurun_out/wide_ncu.log 2>&1
ls -la gpurun_out/prof_widepass_r02.ncu-rep

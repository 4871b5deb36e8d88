#!/bin/bash
# ncu --set full of one HMC launch (2^20 chains, d = 100, L = 20) for the isotropic (packed kernel) and the ill-conditioned
# diagonal Gaussian (config C2's potential)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
cat > /tmp/hmc_one.py <<'PY'
import sys, torch, nfmc_b200
from nfmc_b200.potentials import make_potential
pot = sys.argv[1]
s = nfmc_b200.create_sampler(make_potential(pot, (100,)), event_shape=(100,), strategy="hmc", param_kwargs=dict(n_iterations=10, store_samples=False))
x0 = torch.randn(1 << 20, 100, device="cuda") * 0.1
for _ in range(3):
    out = s.sample(x0, show_progress=False)
torch.cuda.synchronize()
print(pot, out.statistics.acceptance_rate)
PY
for pot in g0 g1; do
  python /tmp/hmc_one.py $pot > gpurun_out/hmc_one_$pot.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:hmc -s 2 -c 1 -o gpurun_out/prof_hmc_${pot}_r02 -f python /tmp/hmc_one.py $pot > gpurun_out/hmc_ncu_$pot.log 2>&1
done
ls -la gpurun_out/prof_hmc_*

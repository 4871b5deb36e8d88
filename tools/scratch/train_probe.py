import os, sys, torch
sys.path.insert(0, "/root/repo")
from nfmc_b200.flow import Flow, RealNVP
d = int(sys.argv[1])
g = torch.Generator().manual_seed(0)
x = torch.randn(4096, d, generator=g).cuda()
torch.manual_seed(1)
f = Flow(RealNVP((d,), n_layers=2)).to("cuda")
f.fit(x, x_val=x, n_epochs=3, lr=0.05, batch_size="adaptive")
torch.cuda.synchronize()

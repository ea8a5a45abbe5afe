"""Exactly N eager train steps of the headline workload (cfg2: batch 32 x 16 x 112 x 112), nothing else -- the program
ncu captures run on (`ncu --metrics ... python tools/gpu_one_step.py 2`): the kernels of a step are the same eager or
replayed from the CUDA graph, and without warm-up / e2e legs a full-step capture stays a few minutes.
    python tools/gpu_one_step.py [steps=2] [batch=32]"""
import os
import sys
import types

os.environ.setdefault("VFD_CUDA_GRAPH", "0")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfd_gan_b200 as V  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
torch.manual_seed(0)
netg, netd = V.NetG(), V.NetD(types.SimpleNamespace(nfr=16, isize=112))
netg.apply(V.weights_init)
netd.apply(V.weights_init)
tr = V.GanTrainStep(netg.to(dev), netd.to(dev), graph=False)
g = torch.Generator().manual_seed(1)
shp3, shp1 = (B, 3, 16, 112, 112), (B, 1, 16, 112, 112)
batch = [(torch.rand(shp3, generator=g) * 2 - 1).to(dev), (torch.rand(shp1, generator=g) > 0.9).float().to(dev),
         (torch.rand(shp3, generator=g) * 2 - 1).to(dev), (torch.rand(shp3, generator=g) * 2 - 1).to(dev)]
for _ in range(steps):
    tr.step(*batch)
torch.cuda.synchronize()
print("losses", tr.losses_dict())

"""Which Python lines issue the torch (non-vfd) ops inside one train step (eager, and while capturing the CUDA graph)."""
import sys, os, types, collections, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
import vfd_gan_b200 as V
from oracle import vfd_oracle as O

WATCH = ("fill_", "zero_", "add_", "add.", "copy_", "mul", "zeros", "clone", "sum", "mean", "_to_copy", "div", "neg", "cat",
         "stack", "expand", "index_put", "foreach")


class Log(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.agg = collections.Counter()

    def __torch_dispatch__(self, func, types_, args=(), kwargs=None):
        name = str(func)
        if "vfd_b200" not in name and any(w in name for w in WATCH):
            fr = [f for f in traceback.extract_stack() if "/vfd_gan_b200/" in f.filename]
            where = " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in fr[-2:][::-1])
            numel = next((a.numel() for a in args if isinstance(a, torch.Tensor)), 0)
            self.agg[(name, where, "big" if numel > 1 << 16 else "small")] += 1
        return func(*args, **(kwargs or {}))


dev = torch.device("cuda", 0)
torch.manual_seed(0)
netg, netd = V.NetG(), V.NetD(types.SimpleNamespace(nfr=16, isize=64))
netg.apply(V.weights_init); netd.apply(V.weights_init)
tr = V.GanTrainStep(netg.to(dev), netd.to(dev), graph=True)
batch = [t.to(dev) for t in O.synthetic_batch(2, 16, 64, seed=0)]
for i in range(4):
    log = Log()
    with log:
        tr.step(*batch)
    torch.cuda.synchronize()
    if i in (1, 2):
        print(f"==== step {i} ({'capture' if i == 2 else 'eager'}): {sum(log.agg.values())} watched aten ops")
        for (name, where, size), c in log.agg.most_common(45):
            print(f"{c:4d} {size:5s} {name:28s} {where}")

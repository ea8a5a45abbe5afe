"""Time video_to_flow at the bench geometry (CUDA events); run under ncu for the per-kernel list.
    python tools/gpu_time_flow.py [B D S reps]"""
import sys
import torch
sys.path.insert(0, ".")
import vfd_gan_b200 as V

B, D, S, reps = (list(map(int, sys.argv[1:5])) + [32, 16, 112, 5][len(sys.argv) - 1:])[:4]
torch.manual_seed(0)
vid = torch.nn.functional.avg_pool3d(torch.rand(B, 3, D, S, S, device="cuda") * 2 - 1, (1, 5, 5), stride=1, padding=(0, 2, 2))
for _ in range(2):
    V.video_to_flow(vid)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    V.video_to_flow(vid)
b.record()
torch.cuda.synchronize()
print(f"video_to_flow {B}x3x{D}x{S}x{S}: {a.elapsed_time(b) / reps:.3f} ms per call")

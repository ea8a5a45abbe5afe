"""Stage-isolation timing of the conv pipelines (vfd_set_debug): for each layer geometry times the forward
(with / without fused BN statistics), dgrad and wgrad kernels with the TMA loads, the MMA issue and the
epilogue switched off in turn. Diagnostics only; prints one table.
    python tools/gpu_stage_probe.py [layer ...]      layer = cin,cout,kd,kh,kw,N,D,H,W"""
import os
import sys
os.environ["VFD_DEBUG_LIB"] = "1"   # route every op through libvfd_b200_debug.so: the switches act on what it launches
import torch
sys.path.insert(0, ".")
from vfd_gan_b200 import ops, _lib

LAYERS = {
    "uconv1.s": (96, 86, 1, 3, 3, 32, 16, 112, 112),
    "uconv1.t": (86, 32, 3, 1, 1, 32, 16, 112, 112),
    "T.dconv1.s": (3, 2, 1, 1, 1, 32, 16, 112, 112),
    "S.dconv1.t": (14, 32, 1, 1, 1, 32, 16, 112, 112),
    "conv_last": (32, 1, 3, 3, 3, 32, 16, 112, 112),
    "T.dconv3.t": (54, 128, 3, 1, 1, 32, 4, 112, 112),
    "uconv2.s": (192, 172, 1, 3, 3, 32, 8, 56, 56),
    "G.dconv1.s": (3, 21, 1, 3, 3, 32, 16, 112, 112),
    "S.dconv1.s": (3, 14, 1, 3, 3, 32, 16, 112, 112),
    "S.dconv2.s": (32, 43, 1, 3, 3, 32, 16, 56, 56),
    "G.dconv1.t": (21, 32, 3, 1, 1, 32, 16, 112, 112),
}
import os
FLAGS = [int(f) for f in os.environ.get("PROBE_FLAGS", "0,1,2,3,4,5,6,7").split(",")]
names = sys.argv[1:] or list(LAYERS)
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
L = _lib.debug_lib()   # the stage switches exist only in libvfd_b200_debug.so
L.vfd_set_debug.argtypes = [_lib._i]


def timeit(fn, reps=3):
    ts = []
    for i in range(reps + 1):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i:
            ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


print(f"{'layer':11s} {'kernel':10s} " + " ".join(f"dbg={d:<6d}" for d in FLAGS) +
      "   (ms; 1=noTMA 2=noMMA 4=noEPI 8=noTMAstore 16=nofence 32=noTMEMld)")
for name in names:
    cin, cout, kd, kh, kw, N, D, H, W = LAYERS[name] if name in LAYERS else map(int, name.split(","))
    cin_p, cout_p = ops.round_up(cin, 8), ops.round_up(cout, 8)
    x = torch.randn(N, D, H, W, cin_p, device=dev).bfloat16()
    x[..., cin:] = 0
    w = torch.randn(cout, cin, kd, kh, kw, device=dev) * 0.05
    pk = ops._packed(w)
    gy = torch.randn(N, D, H, W, cout_p, device=dev).bfloat16()
    gy[..., cout:] = 0
    y = torch.empty(N, D, H, W, cout_p, dtype=torch.bfloat16, device=dev)
    gx = torch.empty_like(x)
    st = torch.zeros(2 * cout_p, dtype=torch.float64, device=dev)
    acc = torch.zeros(kd * kh * kw, cin_p, ops.round_up(cout, 32), dtype=torch.float32, device=dev)
    lay = ops.wgrad_layout(cout, cin, kd, kh, kw, H, W)
    acc3 = torch.zeros(kd * kh * kw, cout_p, ops.round_up(cin, 32), dtype=torch.float32, device=dev)
    kernels = {
        "fwd": lambda: ops.conv3d_fwd(x, pk.fwd, None, y, None, kd, kh, kw, pk.kc_f, cout_p, False),
        "fwd+stats": lambda: ops.conv3d_fwd(x, pk.fwd, None, y, st, kd, kh, kw, pk.kc_f, cout_p, False),
        "dgrad": lambda: ops.conv3d_fwd(gy, pk.dgrad, None, gx, None, kd, kh, kw, pk.kc_d, cin_p, False),
        "wgrad": lambda: ops.conv3d_wgrad(gy, cout, x, cin, acc, kd, kh, kw, False, 0),
        "wgrad3": (lambda: ops.conv3d_wgrad(gy, cout, x, cin, acc3, kd, kh, kw, False, 1)) if lay else None,
    }
    for kname, fn in kernels.items():
        if fn is None:
            continue
        row = []
        for dbg in FLAGS:
            L.vfd_set_debug(dbg)
            row.append(timeit(fn))
        L.vfd_set_debug(0)
        print(f"{name:11s} {kname:10s} " + " ".join(f"{t:10.3f}" for t in row), flush=True)

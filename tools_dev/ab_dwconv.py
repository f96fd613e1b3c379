"""Timing and accuracy of ds_dwconv7 on the level-0..2 shapes of the U-Net (same process, 20 launches each).
DS_LIB_PATH selects an A/B build.
Error: relative L2 against an fp32 torch depthwise convolution of the same 16-bit inputs (first 4 samples)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from diffusynth_b200 import ops
N = 128
for C0, C1, H, W in [(96, 0, 128, 64), (96, 192, 128, 64), (192, 0, 64, 32), (384, 0, 32, 16), (384, 384, 16, 8)]:
    C = C0 + C1
    g = torch.Generator(device="cuda").manual_seed(1)
    s0 = torch.randn((N, H, W, C0), device="cuda", generator=g).to(ops.ACT)
    s1 = torch.randn((N, H, W, C1), device="cuda", generator=g).to(ops.ACT) if C1 else None
    w = (torch.randn((49, C), device="cuda", generator=g) * 0.1)
    tb = torch.randn((1, C + 8), device="cuda", generator=g)
    out = torch.empty((N, H, W, C), dtype=ops.ACT, device="cuda")
    st = ops.dwconv7_stats(N, C, H, W, "cuda")
    for _ in range(3): ops.dwconv7(s0, s1, w, tb, 0, out, N, H, W, stats=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.dwconv7(s0, s1, w, tb, 0, out, N, H, W, stats=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    x = torch.cat([s0[:4]] + ([s1[:4]] if C1 else []), dim=-1).float().permute(0, 3, 1, 2).contiguous()
    ref = F.conv2d(x, w.t().reshape(C, 1, 7, 7), None, padding=3, groups=C) + tb[:, :C, None, None]
    got = out[:4].float().permute(0, 3, 1, 2)
    err = float((got - ref).norm() / ref.norm())
    print(f"dwconv C={C:4d} {H}x{W}: {ms:.4f} ms  {2*49*N*H*W*C/ms/1e9:.1f} TFLOP/s  rel-L2 vs fp32 {err:.2e}", flush=True)

"""One ds_conv_gemm shape in isolation (for ncu).  SHAPE=qkv: to_qkv-like 1x1 96->384; SHAPE=conv1: ConvNeXt conv1 3x3 96->192 with
GroupNorm fold + GELU + statistics (EPI = 2); both at 128x64, N=128."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ops
N, H, W = int(os.environ.get("N", "128")), 128, 64
if os.environ.get("SHAPE", "qkv") == "conv1":
    cin, cout = 96, 192
    w = torch.randn(cout, cin, 3, 3) * (1.0 / (cin * 9) ** 0.5)
    pc = ops.pack_conv_s1(w, torch.randn(cout), 1 + 0.1 * torch.randn(cin), 0.1 * torch.randn(cin)).to("cuda")
    x = torch.randn(N, H, W, cin, device="cuda").to(ops.ACT)
    out = torch.empty((N, H, W, cout), dtype=ops.ACT, device="cuda")
    st = ops.given_stats(torch.zeros(N), torch.ones(N), cin * H * W)
    a, _, keep = ops.conv_args(pc, x, None, N, H, W, out=out, stats_in=st, act=1, want_stats=True)
else:
    cin, cout = 96, 384
    w = torch.randn(cout, cin, 1, 1) * 0.1
    pc = ops.pack_conv_s1(w, None, 1 + 0.1 * torch.randn(cin), 0.1 * torch.randn(cin)).to("cuda")
    x = torch.randn(N, H, W, cin, device="cuda").to(ops.ACT)
    out = torch.empty((N, H, W, cout), dtype=ops.ACT, device="cuda")
    st = ops.given_stats(torch.zeros(N), torch.ones(N), cin * H * W)
    sb = torch.randn(N, cout, device="cuda")
    a, _, keep = ops.conv_args(pc, x, None, N, H, W, out=out, stats_in=st, sbias=sb)
for _ in range(int(os.environ.get("REPS", "3"))):
    ops.run_conv(a)
torch.cuda.synchronize()
print("ok")

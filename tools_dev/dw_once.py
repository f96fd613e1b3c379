"""One ds_dwconv7 launch at the level-0 shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ops
N, C, H, W = 128, 96, 128, 64
s0 = torch.randn((N, H, W, C), device="cuda").to(ops.ACT)
w = torch.randn((49, C), device="cuda") * 0.1
tb = torch.randn((1, C + 8), device="cuda")
out = torch.empty((N, H, W, C), dtype=ops.ACT, device="cuda")
st = ops.dwconv7_stats(N, C, H, W, "cuda")
for _ in range(2):
    ops.dwconv7(s0, None, w, tb, 0, out, N, H, W, stats=st)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))

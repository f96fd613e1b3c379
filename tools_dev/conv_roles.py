"""Per-role wait-cycle breakdown of ds_conv_gemm (DS_CONV_DBG=64 instrumentation) on representative layer shapes.
Needs a library built with the instrumentation: DS_EXTRA_NVCC_FLAGS=-DDS_CONV_DEBUG python -m diffusynth_b200._build --force"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ops

N = int(os.environ.get("N", "128"))
SHAPES = [
    ("l0 conv1 96->192", 96, 0, 192, 3, 128, 64, 1, 0),
    ("l0 conv2 192->96", 192, 0, 96, 3, 128, 64, 0, 1),
    ("l0 conv1 192->192", 96, 96, 192, 3, 128, 64, 1, 0),
    ("l1 conv2 384->192", 384, 0, 192, 3, 64, 32, 0, 1),
    ("l2 conv1 768->768", 384, 384, 768, 3, 32, 16, 1, 0),
    ("l0 to_qkv 96->384", 96, 0, 384, 1, 128, 64, 0, 0),
    ("l0 to_out 128->96", 128, 0, 96, 1, 128, 64, 0, 0),
    ("l0 res_conv 288->96", 96, 192, 96, 1, 128, 64, 0, 0),
    ("l1 to_out 128->192", 128, 0, 192, 1, 64, 32, 0, 0),
]
if os.environ.get("ONLY"):
    SHAPES = [s for s in SHAPES if os.environ["ONLY"] in s[0]]
for name, c0, c1, cout, k, H, W, gelu, res in SHAPES:
    cin = c0 + c1
    w = torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)
    gamma, beta = 1 + 0.1 * torch.randn(cin), 0.1 * torch.randn(cin)
    fold = k == 3
    pc = ops.pack_conv_s1(w, torch.randn(cout), gamma if fold else None, beta if fold else None).to("cuda")
    s0 = torch.randn((N, H, W, c0), device="cuda").to(ops.ACT)
    s1 = torch.randn((N, H, W, c1), device="cuda").to(ops.ACT) if c1 else None
    out = torch.empty((N, H, W, cout), dtype=ops.ACT, device="cuda")
    st = ops.given_stats(torch.zeros(N), torch.ones(N), cin * H * W) if fold else None
    resid = torch.randn((N, H, W, cout), device="cuda").to(ops.ACT) if res else None
    for dbg in os.environ.get("DBGS", "64,95").split(","):
        os.environ["DS_CONV_DBG"] = dbg
        a, so, keep = ops.conv_args(pc, s0, s1, N, H, W, out=out, stats_in=st, act=gelu, residual=resid, want_stats=fold)
        sys.stderr.write(f"{name} dbg={dbg}: "); sys.stderr.flush()
        ops.run_conv(a); ops.run_conv(a)
        torch.cuda.synchronize()

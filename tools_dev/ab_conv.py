"""A/B timing of ds_conv_gemm variants (env switches) on representative U-Net layer shapes, same process, interleaved."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ops

N = int(os.environ.get("N", "128"))
SHAPES = [  # (name, Cin0, Cin1, Cout, k, H, W, gelu, residual)
    ("l0 conv1 96->192", 96, 0, 192, 3, 128, 64, 1, 0),
    ("l0 conv2 192->96", 192, 0, 96, 3, 128, 64, 0, 1),
    ("l0 conv1 288->192", 96, 192, 192, 3, 128, 64, 1, 0),
    ("l0 conv1 192->192", 96, 96, 192, 3, 128, 64, 1, 0),
    ("l1 conv1 576->384", 192, 384, 384, 3, 64, 32, 1, 0),
    ("l1 conv2 384->192", 384, 0, 192, 3, 64, 32, 0, 1),
    ("l2 conv1 768->768", 384, 384, 768, 3, 32, 16, 1, 0),
    ("l0 to_qkv 96->384", 96, 0, 384, 1, 128, 64, 0, 0),
    ("l0 final 96->4", 96, 0, 4, 3, 128, 64, 0, 0),
]
VARIANTS = json.loads(os.environ.get("VARIANTS", '[{"DS_CONV_MAX_SPS": "1"}, {"DS_CONV_MAX_SPS": "4"}]'))


def g(shape, seed):
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    return torch.randn(shape, generator=gen, device="cuda")


for name, c0, c1, cout, k, H, W, gelu, res in SHAPES:
    cin = c0 + c1
    w = torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)
    gamma, beta = 1 + 0.1 * torch.randn(cin), 0.1 * torch.randn(cin)
    fold = k == 3 and cout >= 16
    pc = ops.pack_conv_s1(w, torch.randn(cout), gamma if fold else None, beta if fold else None).to("cuda")
    s0 = g((N, H, W, c0), 1).to(ops.ACT)
    s1 = g((N, H, W, c1), 2).to(ops.ACT) if c1 else None
    out = torch.empty((N, H, W, cout), dtype=ops.ACT, device="cuda") if cout >= 16 else None
    out32 = torch.empty((N, cout, H, W), device="cuda") if cout < 16 else None
    st = ops.given_stats(torch.zeros(N), torch.ones(N), cin * H * W) if fold else None
    resid = g((N, H, W, cout), 3).to(ops.ACT) if res else None
    flops = 2.0 * N * H * W * cout * k * k * cin
    results = []
    for rep in range(3):
        for vi, var in enumerate(VARIANTS):
            for kk, vv in var.items():
                os.environ[kk] = vv
            a, so, keep = ops.conv_args(pc, s0, s1, N, H, W, out=out, out_f32=out32, stats_in=st, act=gelu, residual=resid, want_stats=fold)
            for _ in range(3):
                ops.run_conv(a)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.run_conv(a)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            if rep == 0:
                results.append([])
            results[vi].append(ms)
            for kk in var:
                os.environ.pop(kk, None)
    line = f"{name:20s}"
    for vi, var in enumerate(VARIANTS):
        ms = sorted(results[vi])[1]
        line += f" | {','.join(f'{a}={b}' for a, b in var.items()) or 'default'}: {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF/s"
    print(line, flush=True)

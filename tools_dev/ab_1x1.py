"""Timing of the plain 1x1 convolutions (to_out / res_conv shapes, batch 128) inside a graph of 10; DS_LIB_PATH selects an A/B build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ops
N = 128
for cin, cout, H, W, bn in ((128, 96, 128, 64, 0), (192, 96, 128, 64, 0), (288, 96, 128, 64, 0), (128, 192, 64, 32, 0), (128, 192, 64, 32, 96), (384, 192, 64, 32, 0),
                           (384, 192, 64, 32, 96), (576, 192, 64, 32, 96), (128, 384, 32, 16, 0), (128, 384, 32, 16, 128), (768, 384, 32, 16, 128)):
    w = torch.randn(cout, cin, 1, 1) * cin ** -0.5
    pc = ops.pack_conv_s1(w, torch.randn(cout) * 0.1).to("cuda")
    pc.bn = bn
    x = torch.randn(N, H, W, cin, device="cuda").to(ops.ACT)
    out = torch.empty((N, H, W, cout), dtype=ops.ACT, device="cuda")
    a, st, keep = ops.conv_args(pc, x, None, N, H, W, out=out, want_stats=os.environ.get("STATS", "1") == "1")
    for _ in range(3):
        ops.run_conv(a)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
        for _ in range(10):
            ops.run_conv(a)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    mb = N * H * W * (cin + cout) * 2 / 1e6
    ref = torch.nn.functional.conv2d(x[:2].float().permute(0, 3, 1, 2), w.cuda().to(ops.ACT).float(), pc.e2[0, :cout]).permute(0, 2, 3, 1)
    err = float((out[:2].float() - ref).norm() / ref.norm())
    print(f"{cin:4d}->{cout:4d} BN {bn:3d} {H}x{W}: {ms * 1e3:7.1f} us  {mb / ms / 1e3:6.2f} TB/s  rel err {err:.1e}  mean {float(st.buf[0, 0, 0]) if st is not None else 0.0:+.4f}")

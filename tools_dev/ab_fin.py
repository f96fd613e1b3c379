"""Timing of ds_attn_finalize at the U-Net's attention shapes (batch 128); DS_LIB_PATH selects an A/B build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import _lib, ops
from diffusynth_b200._lib import check
lib = _lib.load()
N = 128
for Cc, H, W in ((96, 128, 64), (192, 64, 32), (384, 32, 16), (384, 16, 8)):
    npix = H * W
    part = torch.rand((lib.ds_attn_part_floats(N, 4, npix),), device="cuda")
    wout = torch.randn((Cc, 128), device="cuda")
    M = torch.empty((N, Cc, 128), dtype=ops.ACT, device="cuda")
    run = lambda: check(lib.ds_attn_finalize(part.data_ptr(), wout.data_ptr(), M.data_ptr(), N, 4, npix, Cc, Cc, ops._stream()), "fin")
    for _ in range(3):
        run()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
        for _ in range(20):
            run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"C={Cc:4d} {H}x{W}: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us per finalize (graph of 20)  checksum {float(M.float().abs().mean()):.6f}")

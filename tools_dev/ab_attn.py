"""Timing of ds_attn_qkv_ctx at the U-Net's attention shapes (batch 128); DS_LIB_PATH selects an A/B build (AQ_DBG switches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import _lib, ops
from diffusynth_b200._lib import check

lib = _lib.load()
N = 128
for Cc, H, W in ((96, 128, 64), (192, 64, 32), (384, 32, 16), (384, 16, 8)):
    npix = H * W
    x = torch.randn((N, H, W, Cc), device="cuda").to(ops.ACT)
    wq = torch.randn(384, Cc, 1, 1) * (2.0 / Cc ** 0.5)
    pc = ops.pack_conv_s1(wq, None, torch.ones(Cc), torch.zeros(Cc)).to("cuda")
    st = ops.given_stats(torch.zeros(N), torch.ones(N), npix * Cc)
    sb = torch.zeros((N, 384), device="cuda")
    qp = torch.empty((N, H, W, 128), dtype=ops.ACT, device="cuda")
    part = torch.zeros((lib.ds_attn_part_floats(N, 4, npix),), device="cuda")
    def run():
        check(lib.ds_attn_qkv_ctx(x.data_ptr(), Cc, 0, st.buf.data_ptr(), st.slots, pc.weight.data_ptr(), pc.e1.data_ptr(), pc.e2.data_ptr(),
                                  sb.data_ptr(), sb.stride(0), qp.data_ptr(), part.data_ptr(), N, 4, npix, 32 ** -0.5, ops._stream()), "qkv_ctx")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"C={Cc:4d} {H}x{W}: {ms:.4f} ms  ({2.0 * N * npix * Cc * 384 / ms / 1e9:.0f} TFLOP/s GEMM)")

"""ds_attn_qkv_ctx (to_qkv GEMM + q soft-max + partial context in one tcgen05 kernel, k / v never written) against the separate
ds_conv_gemm(to_qkv) + ds_attn_ctx_partial launches on the same inputs, and against an fp32 torch evaluation of
LinearCrossAttentionAdd's core (diffusion_components.py:271-289): q' and the folded per-sample matrix M = Wout . ctx^T."""
import ctypes as C

import pytest
import torch

from diffusynth_b200 import _lib, ops
from diffusynth_b200._lib import check
from oracle import cases
from tests.gpu_util import nhwc, rel

pytestmark = pytest.mark.gpu
HID, HEADS, DH = 128, 4, 32


@pytest.mark.parametrize("N,Cc,H,W,mod", [(3, 96, 24, 20, 0), (2, 192, 40, 32, 0), (4, 96, 32, 16, 2), (2, 384, 16, 8, 0), (2, 64, 6, 5, 0)])
def test_fused_qkv_ctx_matches_separate_kernels(N, Cc, H, W, mod):
    lib = _lib.load()
    npix = H * W
    n_in = mod if mod else N
    x = cases.randn((n_in, Cc, H, W), 1) * 1.7 + 0.3
    wq = cases.randn((3 * HID, Cc, 1, 1), 2) * (2.0 / Cc ** 0.5)
    gamma, beta = 1.0 + 0.2 * cases.randn((Cc,), 3), 0.2 * cases.randn((Cc,), 4)
    sb = cases.randn((N, 3 * HID), 5) * 0.5
    sb[:, 2 * HID:] = 0
    wout = (cases.randn((Cc, HID), 6) * 0.1).cuda().contiguous()
    xs = nhwc(x)
    mean, var = x.mean(dim=(1, 2, 3)), x.var(dim=(1, 2, 3), unbiased=False)
    st = ops.given_stats(mean, (var + 1e-5).rsqrt(), npix * Cc)
    pc = ops.pack_conv_s1(wq, None, gamma, beta).to("cuda")
    sbd = sb.cuda().contiguous()
    scale = DH ** -0.5
    stream = ops._stream()

    def finalize(part):
        M = torch.empty((N, ops.pad16(Cc), HID), dtype=ops.ACT, device="cuda")
        check(lib.ds_attn_finalize(part.data_ptr(), wout.data_ptr(), M.data_ptr(), N, HEADS, npix, Cc, ops.pad16(Cc), stream), "fin")
        return M.float().cpu()

    # separate kernels
    qkv = torch.zeros((N, H, W, 3 * HID), dtype=ops.ACT, device="cuda")
    a, _, keep = ops.conv_args(pc, xs, None, N, H, W, out=qkv, stats_in=st, sbias=sbd, src_batch_mod=mod)
    ops.run_conv(a)
    qp1 = torch.zeros((N, H, W, HID), dtype=ops.ACT, device="cuda")
    part1 = torch.zeros((lib.ds_attn_part_floats(N, HEADS, npix),), dtype=torch.float32, device="cuda")
    check(lib.ds_attn_ctx_partial(qkv.data_ptr(), qp1.data_ptr(), part1.data_ptr(), N, HEADS, npix, 0, scale, stream), "ctx")
    M1 = finalize(part1)
    # fused kernel
    qp2 = torch.zeros((N, H, W, HID), dtype=ops.ACT, device="cuda")
    part2 = torch.zeros_like(part1)
    check(lib.ds_attn_qkv_ctx(xs.data_ptr(), Cc, mod, st.buf.data_ptr(), st.slots, pc.weight.data_ptr(), pc.e1.data_ptr(), pc.e2.data_ptr(),
                              sbd.data_ptr(), sbd.stride(0), qp2.data_ptr(), part2.data_ptr(), N, HEADS, npix, scale, stream), "qkv_ctx")
    M2 = finalize(part2)
    torch.cuda.synchronize()
    # fp32 torch evaluation of the same math on the 16-bit-rounded inputs
    xr = xs.float().cpu().permute(0, 3, 1, 2)
    if mod:
        xr = xr.repeat(N // mod, 1, 1, 1)
    xn = torch.nn.functional.group_norm(xr, 1, gamma, beta, eps=1e-5)
    qkv_ref = torch.nn.functional.conv2d(xn, wq) + sb.view(N, -1, 1, 1)
    q, k, v = [t.reshape(N, HEADS, DH, npix) for t in qkv_ref.chunk(3, dim=1)]
    qs = q.softmax(dim=2) * scale
    ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(dim=3), v)
    M_ref = torch.einsum("chE,bhdE->bchd", wout.cpu().view(Cc, HEADS, DH), ctx).reshape(N, Cc, HID)
    q_ref = qs.reshape(N, HID, H, W).permute(0, 2, 3, 1)
    e_q12, e_q2, e_M12, e_M1, e_M2 = rel(qp2, qp1), rel(qp2, q_ref), rel(M2, M1), rel(M1[:, :Cc], M_ref), rel(M2[:, :Cc], M_ref)
    print(f"\nN={N} C={Cc} {H}x{W}: q' fused vs separate {e_q12:.2e}, vs fp32 {e_q2:.2e}; M fused vs separate {e_M12:.2e}; vs fp32: separate {e_M1:.2e}, fused {e_M2:.2e}")
    assert e_q2 < 3e-3 and e_M2 < 5e-3 and e_M12 < 5e-3

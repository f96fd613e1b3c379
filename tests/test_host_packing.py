"""Host logic without a GPU: the weight packing / tap tables / GroupNorm folding that feed ds_conv_gemm
are checked by evaluating the kernel's documented contract (include/diffusynth_b200.h) in torch on the CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffusynth_b200 import ops
from oracle import cases


def contract(pc, srcs, kind, mean_rstd=None, act=0, residual=None):
    """ds_conv_gemm's contract restated: srcs = list of NCHW fp32 tensors (channel-concatenated)."""
    x = torch.cat(srcs, dim=1).to(ops.ACT).float()
    N, C, Hin, Win = x.shape
    w = pc.weight.float()                                    # [G, Cout_pad, ntaps*C]
    G, ntaps = len(pc.taps), len(pc.taps[0])
    if kind == "down":
        Hg, Wg = Hin // 2, Win // 2
        views = [x[:, :, py::2, px::2] for py in range(2) for px in range(2)]
    else:
        Hg, Wg, views = Hin, Win, [x]
    outs = []
    for g in range(G):
        acc = torch.zeros(N, pc.cout_pad, Hg, Wg)
        for t, (dy, dx, v) in enumerate(pc.taps[g]):
            src = F.pad(views[v], (8, 8, 8, 8))[:, :, 8 + dy:8 + dy + Hg, 8 + dx:8 + dx + Wg]
            acc += torch.einsum("oc,nchw->nohw", w[g, :, t * C:(t + 1) * C], src)
        outs.append(acc)
    hh = torch.arange(Hg).view(-1, 1).expand(Hg, Wg)
    ww = torch.arange(Wg).view(1, -1).expand(Hg, Wg)
    cls = torch.zeros(Hg, Wg, dtype=torch.long)
    if pc.ncls == 9:
        cls = torch.where(hh == 0, 0, torch.where(hh == Hg - 1, 2, 1)) * 3 + torch.where(ww == 0, 0, torch.where(ww == Wg - 1, 2, 1))
    res = []
    for acc in outs:
        mean, rstd = (0.0, 1.0) if mean_rstd is None else mean_rstd
        v = acc * rstd + pc.e2[cls].permute(2, 0, 1)[None]
        if pc.e1 is not None and mean_rstd is not None:
            v = v - (mean * rstd) * pc.e1[cls].permute(2, 0, 1)[None]
        if act == 1:
            v = F.gelu(v)
        res.append(v[:, :pc.cout])
    if kind == "up":
        out = torch.zeros(N, pc.cout, 2 * Hg, 2 * Wg)
        for g in range(4):
            out[:, :, g // 2::2, g % 2::2] = res[g]
    else:
        out = res[0]
    return out + (residual if residual is not None else 0)


def rel(a, b):
    return float((a - b).norm() / b.norm())


def test_conv3x3_with_folded_groupnorm():
    x0, x1 = cases.randn((2, 32, 8, 16), 1) + 0.7, cases.randn((2, 64, 8, 16), 2) * 2
    w, b = cases.randn((48, 96, 3, 3), 3) * 0.05, cases.randn((48,), 4)
    gamma, beta = 1 + 0.2 * cases.randn((96,), 5), 0.3 * cases.randn((96,), 6)
    pc = ops.pack_conv_s1(w, b, gamma, beta)
    xb = torch.cat([x0, x1], 1).to(ops.ACT).float()
    out = torch.empty(2, 48, 8, 16)
    ref = torch.empty_like(out)
    for n in range(2):       # GroupNorm(1, C): per-sample scalars
        mean, var = xb[n].mean(), xb[n].var(unbiased=False)
        rstd = (var + 1e-5).rsqrt()
        out[n] = contract(pc, [x0[n:n + 1], x1[n:n + 1]], "s1", (mean, rstd))[0]
        ref[n] = F.conv2d(F.group_norm(xb[n:n + 1], 1, gamma, beta, 1e-5), w, b, padding=1)[0]
    assert rel(out, ref) < 4e-3      # only the bf16 rounding of gamma*W separates them


@pytest.mark.parametrize("k", [1, 3])
def test_plain_conv(k):
    x = cases.randn((1, 64, 6, 8), 7)
    w, b = cases.randn((20, 64, k, k), 8) * 0.1, cases.randn((20,), 9)
    pc = ops.pack_conv_s1(w, b)
    assert pc.cout_pad == 32 and pc.e1 is None
    ref = F.conv2d(x.to(ops.ACT).float(), w.to(ops.ACT).float(), b, padding=k // 2)
    assert rel(contract(pc, [x], "s1"), ref) < 1e-5


def test_downsample_as_parity_views():
    x = cases.randn((2, 32, 8, 12), 10)
    w, b = cases.randn((32, 32, 4, 4), 11) * 0.1, cases.randn((32,), 12)
    pc = ops.pack_conv_down(w, b)
    ref = F.conv2d(x.to(ops.ACT).float(), w.to(ops.ACT).float(), b, stride=2, padding=1)
    assert rel(contract(pc, [x], "down"), ref) < 1e-5


def test_transposed_conv_as_four_phases():
    x = cases.randn((2, 32, 5, 6), 13)
    w, b = cases.randn((32, 48, 4, 4), 14) * 0.1, cases.randn((48,), 15)
    pc = ops.pack_conv_up(w, b)
    ref = F.conv_transpose2d(x.to(ops.ACT).float(), w.to(ops.ACT).float(), b, stride=2, padding=1)
    assert rel(contract(pc, [x], "up"), ref) < 1e-5


def test_tile_and_blocking_choices():
    assert ops.choose_tile(128, 64) == (2, 64) and ops.choose_tile(16, 8) == (16, 8) and ops.choose_tile(128, 24) == (16, 8)
    assert ops.choose_bn(768) == 256 and ops.choose_bn(384) == 192 and ops.choose_bn(96) == 96 and ops.choose_bn(16) == 16
    # jobs too small to fill the SMs with the widest tiling take narrow N tiles (one-prompt sampling, deep levels of small batches);
    # the batch-64 job (>= 128 M-tiles on every level) never does
    assert ops.choose_bn(384, 2) == 64 and ops.choose_bn(768, 2) == 64 and ops.choose_bn(96, 32) == 32 and ops.choose_bn(16, 2) == 16
    assert ops.choose_bn(384, 128) == 192 and ops.choose_bn(96, 128) == 96 and ops.choose_bn(192, 80) == 192 and ops.choose_bn(192, 32) == 64
    for hw in [(128, 64), (64, 32), (32, 16), (16, 8), (512, 256), (128, 144), (128, 24)]:
        hb, wb = ops.choose_tile(*hw)
        assert hb * wb == 128


@pytest.mark.parametrize("H,W", [(16, 7), (9, 5), (32, 13), (8, 8)])
def test_conv_args_odd_sizes_and_padded_upsample(H, W):
    """Host side of pad_to_match (diffusion_components.py:210-232) without a GPU: the stride-2 conv's four parity views get their
    own extents when the input size is odd, and the ConvTranspose's output addressing (per-phase offsets + strides) lands the
    upsampled map at (delta // 2) inside a larger zero map -- checked by replaying the addressing on the CPU."""
    N, C = 2, 32
    x = torch.zeros((N, H, W, C), dtype=ops.ACT)
    pc = ops.pack_conv_down(torch.zeros(C, C, 4, 4), torch.zeros(C))
    out = torch.zeros((N, H // 2, W // 2, C), dtype=ops.ACT)
    a, _, _ = ops.conv_args(pc, x, None, N, H, W, out=out)
    assert (a.H, a.W) == (H // 2, W // 2) and a.num_views == 4
    for v in range(4):
        py, px = v // 2, v % 2
        assert (a.view_hv[v], a.view_wv[v]) == (len(range(py, H, 2)), len(range(px, W, 2)))
        assert a.view_off[v] == py * W + px
    # the last tap column of the last output pixel exists exactly when the even-parity view is one wider
    assert (2 * (W // 2 - 1) + 2 < W) == (a.view_wv[0] > W // 2)
    # up: (H//2, W//2) -> (2*(H//2), 2*(W//2)) written into an (H, W) map
    pcu = ops.pack_conv_up(torch.zeros(C, C, 4, 4), torch.zeros(C))
    y = torch.zeros((N, H // 2, W // 2, C), dtype=ops.ACT)
    big = torch.zeros((N, H, W, C), dtype=ops.ACT)
    au, _, _ = ops.conv_args(pcu, y, None, N, H // 2, W // 2, out=big)
    hit = torch.zeros(N * H * W * C, dtype=torch.int32)
    for g in range(4):
        for n in range(N):
            for h in range(H // 2):
                for w in range(W // 2):
                    off = au.out_goff[g] + n * au.out_sn + h * au.out_sh + w * au.out_sw
                    hit[off:off + C] += 1
    hit = hit.view(N, H, W, C)
    Ho, Wo = 2 * (H // 2), 2 * (W // 2)
    top, left = (H - Ho) // 2, (W - Wo) // 2
    expect = torch.zeros(N, H, W, C, dtype=torch.int32)
    expect[:, top:top + Ho, left:left + Wo] = 1
    assert torch.equal(hit, expect)          # every upsampled pixel written exactly once, the padding never

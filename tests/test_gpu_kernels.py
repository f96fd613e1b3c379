"""Bandwidth-bound kernels through the C ABI against the CPU oracle on the same seeded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffusynth_b200 import weights as W
from oracle import cases, ds_oracle as O
from tests.gpu_util import EPS16, bf, nchw, nhwc, rel
from diffusynth_b200 import ops

pytestmark = pytest.mark.gpu


def lib():
    from diffusynth_b200 import _lib
    return _lib.load()


def S():
    return torch.cuda.current_stream().cuda_stream


def test_device_is_blackwell():
    assert lib().ds_check_device(0) == 0 and lib().ds_version() >= 100


@pytest.mark.parametrize("eta", [0.0, 1.0])
def test_ddim_step_matches_reference_formula(eta):
    from diffusynth_b200 import ops
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, 20, dtype=np.int32)))
    x, eu, ec, z = (cases.randn((3, 4, 128, 64), s) for s in (1, 2, 3, 4))
    for t in (19, 7, 0):
        co = O.ddim_coefficients(sch, t, eta)
        ref = O.ddim_update(x, eu, ec, 6, co, z)
        coef = torch.tensor([co["sqrt_one_minus_at"], co["sqrt_at"], co["sqrt_ap"], co["dir_coef"], co["sigma"], 6.0, 0, 0], dtype=torch.float32).cuda()
        out = torch.empty_like(x).cuda()
        ops.ddim_step(eu.cuda(), ec.cuda(), x.cuda(), z.cuda(), coef, out)
        assert rel(out, ref) < 1e-6          # fp32, same operation order: differences are last-bit only
        ref1 = O.ddim_update(x, None, ec, 1.0, co, z)
        ops.ddim_step(None, ec.cuda(), x.cuda(), z.cuda() if eta else None, coef, out)
        assert rel(out, ref1) < 1e-6


def test_sampler_coefficients_match_oracle():
    from diffusynth_b200.sampler import DiffSynthSampler
    s = DiffSynthSampler(1000, device="cuda", mute=True)
    s.respace(list(np.linspace(0, 999, 20, dtype=np.int32)))
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, 20, dtype=np.int32)))
    assert s.timestep_map == sch.timestep_map and np.array_equal(s.alphas_cumprod, sch.alphas_cumprod)
    for eta in (0.0, 1.0):
        for t in range(20):
            co = O.ddim_coefficients(sch, t, eta)
            mine = s._coef(t, eta)
            assert mine[:5] == [co["sqrt_one_minus_at"], co["sqrt_at"], co["sqrt_ap"], co["dir_coef"], co["sigma"]]


def test_q_sample_and_mask_blend():
    from diffusynth_b200 import ops
    sch = O.Schedule(1000)
    x0, nz, img = (cases.randn((2, 4, 128, 40), s) for s in (5, 6, 7))
    mask = (cases.randn((2, 1, 128, 40), 8) > 0).float()
    ref = O.q_sample(sch, x0, 500, nz)
    coef = torch.tensor([np.float32(sch.sqrt_alphas_cumprod[500]), np.float32(sch.sqrt_one_minus_alphas_cumprod[500])], dtype=torch.float32).cuda()
    out = torch.empty_like(x0).cuda()
    ops.q_sample(x0.cuda(), nz.cuda(), coef, out)
    assert rel(out, ref) < 1e-7
    blended = img.clone().cuda()
    ops.mask_blend(x0.cuda(), nz.cuda(), mask.cuda(), coef, blended)
    assert rel(blended, mask * ref + (1 - mask) * img) < 1e-7


@pytest.mark.parametrize("C0,C1,H,W,N", [(96, 0, 128, 64, 2), (96, 192, 32, 16, 2), (384, 384, 16, 8, 3), (32, 0, 20, 24, 1)])
def test_dwconv7(C0, C1, H, W, N):
    from diffusynth_b200 import ops
    Cc = C0 + C1
    x = cases.randn((N, Cc, H, W), 9)
    w, tb = cases.randn((Cc, 1, 7, 7), 10) * 0.1, cases.randn((N, Cc + 5), 11)
    out = torch.zeros((N, H, W, Cc), dtype=ops.ACT, device="cuda")
    st = ops.dwconv7_stats(N, Cc, H, W, "cuda")
    s0 = nhwc(x[:, :C0])
    s1 = nhwc(x[:, C0:]) if C1 else None
    ops.dwconv7(s0, s1, w.reshape(Cc, 49).t().contiguous().cuda(), tb.cuda(), tb.shape[1], out, N, H, W, stats=st)
    ref = F.conv2d(bf(x), w, None, padding=3, groups=Cc) + tb[:, :Cc, None, None]
    assert rel(nchw(out), ref) < EPS16
    buf = st.buf.double().cpu()
    s = buf[:, 2:].sum(dim=1)
    assert torch.allclose(s[:, 0], ref.double().sum(dim=(1, 2, 3)), rtol=1e-3, atol=1.0)
    assert torch.allclose(s[:, 1], (ref.double() ** 2).sum(dim=(1, 2, 3)), rtol=1e-3)
    assert torch.allclose(buf[:, 0, 0], ref.double().mean(dim=(1, 2, 3)), rtol=1e-3, atol=1e-4)
    assert torch.allclose(buf[:, 0, 1], (ref.double().var(dim=(1, 2, 3), unbiased=False) + 1e-5).rsqrt(), rtol=1e-3)


def test_stem_im2col():
    """init_conv as im2col + GEMM (diffusion.py:82): col[n, h, w, ky*32 + kx*4 + ci] = x[n, ci, h + ky - 3, w + kx - 3] (zero padded,
    8th pixel slot of every kernel row zero); the GEMM itself is covered by the init_conv tap of tests/test_gpu_model.py."""
    x = cases.randn((2, 4, 32, 24), 12) * 3
    col = torch.full((2, 32, 24, 224), 7.0, dtype=ops.ACT, device="cuda")
    assert lib().ds_stem_im2col(x.cuda().data_ptr(), col.data_ptr(), 2, 4, 32, 24, S()) == 0
    xp = F.pad(x, (3, 3, 3, 3))
    ref = torch.zeros((2, 32, 24, 7, 8, 4))
    for ky in range(7):
        for kx in range(7):
            ref[:, :, :, ky, kx, :] = xp[:, :, ky:ky + 32, kx:kx + 24].permute(0, 2, 3, 1)
    assert torch.equal(col.float().cpu(), ref.reshape(2, 32, 24, 224).to(ops.ACT).float())


def test_time_and_condition_linears():
    from diffusynth_b200 import ops
    sd = W.unet_random_state_dict(seed=0)
    t = torch.tensor([947, 52, 0, 999], dtype=torch.long)
    ref = O.time_embedding(sd, t, 96)
    sin = torch.empty((4, 96), device="cuda")
    assert lib().ds_sinusoidal_embedding(t.cuda().data_ptr(), sin.data_ptr(), 4, 96, S()) == 0
    t1, te = torch.empty((4, 384), device="cuda"), torch.empty((4, 384), device="cuda")
    ops.linear(sin, sd["time_mlp.1.weight"].cuda(), sd["time_mlp.1.bias"].cuda(), t1, act_out=1)
    ops.linear(t1, sd["time_mlp.3.weight"].cuda(), sd["time_mlp.3.bias"].cuda(), te)
    assert rel(te, ref) < 2e-4       # sin/cos of arguments up to ~1000 rad in fp32


def _attn_ref(qkv, heads, softmax_q):
    B, n, _ = qkv.shape
    q, k, v = (qkv[..., i * heads * 32:(i + 1) * heads * 32].reshape(B, n, heads, 32).permute(0, 2, 3, 1) for i in range(3))
    if softmax_q:
        q = q.softmax(dim=-2) * 32 ** -0.5
    ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(dim=-1), v)
    return q, ctx


@pytest.mark.parametrize("heads,n,mode", [(4, 8192, 0), (4, 128, 0), (1, 2048 + 40, 1)])
def test_linear_attention_core(heads, n, mode):
    B, hid, Cc = 2, heads * 32, 96
    qkv = cases.randn((B, n, 3 * hid), 15) * 1.5
    wout = cases.randn((Cc, hid), 16) * 0.1
    qd = qkv.to(ops.ACT).cuda()
    qp = torch.zeros((B, n, hid), dtype=ops.ACT, device="cuda")
    part = torch.empty((lib().ds_attn_part_floats(B, heads, n),), device="cuda")
    M = torch.zeros((B, Cc, hid), dtype=ops.ACT, device="cuda")
    assert lib().ds_attn_ctx_partial(qd.data_ptr(), qp.data_ptr(), part.data_ptr(), B, heads, n, mode, C.c_float(32 ** -0.5), S()) == 0
    assert lib().ds_attn_finalize(part.data_ptr(), wout.cuda().data_ptr(), M.data_ptr(), B, heads, n, Cc, Cc, S()) == 0
    q_ref, ctx = _attn_ref(bf(qkv), heads, mode == 0)
    assert rel(qp.float().cpu().reshape(B, n, heads, 32).permute(0, 2, 3, 1), q_ref) < EPS16
    M_ref = torch.einsum("che,bhde->bchd", wout.view(Cc, heads, 32), ctx).reshape(B, Cc, hid)
    assert rel(M, M_ref) < EPS16


def test_gn_apply_residual():
    N, H, Wd, Cc = 2, 16, 8, 96
    y, x = cases.randn((N, Cc, H, Wd), 17) * 2 + 1, cases.randn((N, Cc, H, Wd), 18)
    g, b = 1 + 0.1 * cases.randn((Cc,), 19), 0.1 * cases.randn((Cc,), 20)
    yb = bf(y)
    st = ops.given_stats(yb.mean(dim=(1, 2, 3)), (yb.var(dim=(1, 2, 3), unbiased=False) + 1e-5).rsqrt(), Cc * H * Wd)
    out = torch.zeros((N, H, Wd, Cc), dtype=ops.ACT, device="cuda")
    yd, xd = nhwc(y), nhwc(x)
    assert lib().ds_gn_apply_residual(yd.data_ptr(), xd.data_ptr(), out.data_ptr(), st.buf.data_ptr(), st.slots,
                                      g.cuda().data_ptr(), b.cuda().data_ptr(), N, Cc, H * Wd, 0, S()) == 0
    assert rel(nchw(out), F.group_norm(yb, 1, g, b, 1e-5) + bf(x)) < EPS16
    # residual shared by both halves of a guidance-doubled batch: sample n adds x[n % 1]
    assert lib().ds_gn_apply_residual(yd.data_ptr(), xd.data_ptr(), out.data_ptr(), st.buf.data_ptr(), st.slots,
                                      g.cuda().data_ptr(), b.cuda().data_ptr(), N, Cc, H * Wd, 1, S()) == 0
    assert rel(nchw(out), F.group_norm(yb, 1, g, b, 1e-5) + bf(x)[:1]) < EPS16


@pytest.mark.parametrize("act", [1, 2])
def test_group_norm16_act(act):
    N, H, Wd, Cc, Cp, G = 2, 32, 16, 80, 96, 16
    x = cases.randn((N, Cc, H, Wd), 21) * 2 + 0.5
    g, b = 1 + 0.1 * cases.randn((Cc,), 22), 0.1 * cases.randn((Cc,), 23)
    xd = nhwc(x, Cp)
    part = torch.empty((N, G, 64, 2), device="cuda")
    out = torch.zeros_like(xd)
    assert lib().ds_group_stats(xd.data_ptr(), part.data_ptr(), N, Cc, Cp, G, H * Wd, 64, S()) == 0
    assert lib().ds_gn_act(xd.data_ptr(), out.data_ptr(), part.data_ptr(), 64, g.cuda().data_ptr(), b.cuda().data_ptr(), N, Cc, Cp, G, H * Wd,
                           C.c_float(1e-6), act, S()) == 0
    r = F.group_norm(bf(x), G, g, b, 1e-6)
    r = F.relu(r) if act == 1 else r * torch.sigmoid(r)
    assert rel(nchw(out, Cc), r) < EPS16 and float(out[..., Cc:].abs().max()) == 0.0


def test_vq_indices_bit_exact(golden):
    sd = W.vqgan_random_state_dict(seed=1)
    cb = sd["_vq_vae._embedding.weight"]
    for lat in (cases.vq_latents(), cases.randn((3, 4, 128, 24), 77) * 2):
        q_ref, idx_ref = O.vq_quantize(lat, cb)
        x = lat.cuda()
        out = torch.empty_like(x)
        idx = torch.empty((lat.numel() // 4,), dtype=torch.long, device="cuda")
        assert lib().ds_vq_quantize(x.data_ptr(), cb.cuda().data_ptr(), 8192, out.data_ptr(), idx.data_ptr(), lat.shape[0], lat.shape[2] * lat.shape[3], S()) == 0
        assert torch.equal(idx.cpu(), idx_ref), int((idx.cpu() != idx_ref).sum())
        assert torch.equal(out.cpu(), q_ref)
    assert np.array_equal(O.vq_quantize(cases.vq_latents(), cb)[1].numpy()[:100], golden["vqgan"]["vq_idx"][:100].astype(np.int64))


def test_istft_and_stft_kernels():
    from diffusynth_b200 import codec
    spec = cases.randn((2, 3, 512, 256), 24)
    spec[:, 0] = spec[:, 0].abs() * 0.8
    wave = codec.spectrogram_to_waveform(spec.cuda()).cpu()
    assert tuple(wave.shape) == (2, 65280)
    for b in range(2):
        ref = O.spectrogram_to_waveform(spec[b].numpy().astype(np.float64))
        assert rel(wave[b], torch.from_numpy(ref)) < 1e-5
    short = codec.spectrogram_to_waveform(spec[:, :, :, :37].contiguous().cuda()).cpu()      # ragged frame count
    assert rel(short[1], torch.from_numpy(O.spectrogram_to_waveform(spec[1, :, :, :37].numpy().astype(np.float64)))) < 1e-5
    w = torch.from_numpy(np.stack([cases.synthetic_wave(seed=31), cases.synthetic_wave(seed=32)])).float()
    enc = codec.waveform_to_spectrogram(w.cuda()).cpu()
    for b in range(2):
        ref = O.waveform_to_spectrogram(w[b].numpy().astype(np.float64))
        assert rel(enc[b, 0], torch.from_numpy(ref[0])) < 1e-4
        strong = torch.from_numpy(ref[0]) > 1e-2          # phase is ill-conditioned where the magnitude vanishes
        assert float((enc[b, 1] - torch.from_numpy(ref[1]))[strong].abs().max()) < 5e-3
        assert float((enc[b, 2] - torch.from_numpy(ref[2]))[strong].abs().max()) < 5e-3
    # encode -> decode round trip at full size
    back = codec.spectrogram_to_waveform(enc.cuda()).cpu()
    assert float((back[:, 2048:-2048] - (w - w.mean(dim=1, keepdim=True))[:, 2048:-2048]).abs().max()) < 5e-3


def _circ(a, b):
    """Distance between uint8 images modulo 256 (the phase image wraps where numpy's cast does)."""
    d = (a.astype(np.int16) - b.astype(np.int16)) % 256
    return np.minimum(d, 256 - d)


def test_decode_glue_images_and_lists(golden):
    """ds_spec_images / ds_latent_image and the 6-list decode glue vs the reference's own output (tests/golden/extra.npz)
    and the oracle.  The latent image is bit-exact (fp32 sub/div/mul); the spectrogram / phase images pass through
    expm1f / float64 log / atan2, so a pixel may sit one level off where the value lands within float round-off of an
    integer: at most 1 level, on fewer than 0.5 % of the pixels."""
    from diffusynth_b200 import codec
    g = golden["extra"]
    spec, other = cases.spec_representation(), cases.spec_representation(seed=53)

    class FakeDecoder:
        def __call__(self, z):
            return spec.cuda()

    out = codec.encodeBatch2GradioOutput_STFT(FakeDecoder(), torch.zeros(2, 4, 128, 3), resolution=(512, 12), original_STFT_batch=other.numpy())
    assert [len(o) for o in out] == [2] * 6
    for key, lst in zip(("mag_img", "phase_img", "mag_img_amp", "phase_img_amp"), (out[0], out[1], out[3], out[4])):
        got, ref = np.stack(lst), g[f"glue_{key}"]
        assert got.dtype == np.uint8 and got.shape == ref.shape == (2, 513, 12, 3)
        d = _circ(got, ref)
        assert d.max() <= 1 and (d != 0).mean() < 5e-3, (key, d.max(), (d != 0).mean())
    for key, lst in zip(("signal", "signal_amp"), (out[2], out[5])):
        assert rel(torch.from_numpy(np.stack(lst)), torch.from_numpy(g[f"glue_{key}"])) < 1e-5
    assert codec.encodeBatch2GradioOutput_STFT(FakeDecoder(), torch.zeros(2, 4, 128, 3))[3:] == ([], [], [])
    lat = cases.small_latents()
    img = codec.latent_images(lat.cuda()).cpu().numpy()
    assert np.array_equal(img, g["latent_img"])
    assert np.array_equal(codec.latent_representation_to_Gradio_image(lat[1]), g["latent_img"][1])
    # full size (batch 8 of [3,512,256]) against the oracle on one sample; a constant channel gives NaN -> 0 like numpy
    big = cases.spec_representation(B=8, T=256, seed=54)
    mag, ph = codec.spectrogram_images(big.cuda())
    m_ref, p_ref, _ = O.decode_products(big[5].numpy())
    for got, ref in ((mag[5].cpu().numpy(), m_ref), (ph[5].cpu().numpy(), p_ref)):
        d = _circ(got, ref)
        assert d.max() <= 1 and (d != 0).mean() < 5e-3
    flat = torch.zeros((1, 4, 16, 8)); flat[0, 1] = cases.randn((16, 8), 55)
    assert np.array_equal(codec.latent_images(flat.cuda())[0].cpu().numpy(), O.latent_image(flat[0].numpy()))


def test_input_batch_encode_glue():
    """InputBatch2Encode_STFT (utils.py:131-191): latents / quantised latents from the VQGAN encoder + quantiser, and the
    renderings of the INPUT representation."""
    from diffusynth_b200 import VQGAN, codec
    vq = VQGAN(**W.VQGAN_DEPLOYED, device="cuda")
    vq.load_state_dict(W.vqgan_random_state_dict(seed=1))
    spec = torch.from_numpy(O.waveform_to_spectrogram(cases.synthetic_wave())[None])
    imgs, phases, signals, lat, q = codec.InputBatch2Encode_STFT(vq._encoder, spec, quantizer=vq._vq_vae)
    assert tuple(lat.shape) == tuple(q.shape) == (1, 4, 128, 64) and len(imgs) == len(phases) == len(signals) == 1
    m_ref, p_ref, s_ref = O.decode_products(spec[0].numpy())
    d = _circ(imgs[0], m_ref)
    assert d.max() <= 1 and (d != 0).mean() < 5e-3
    assert rel(torch.from_numpy(signals[0]), torch.from_numpy(s_ref)) < 1e-5
    # quantizer=None is the reference's VAE branch (utils.py:162-164): the encoder returns (mu, logvar, z), no quantised batch
    class _VaeEncoder:
        def parameters(self):
            return vq._encoder.parameters()

        def __call__(self, x):
            z = vq._encoder(x)
            return z * 0.5, z * 0.0, z

    imgs2, _, signals2, lat2, q2 = codec.InputBatch2Encode_STFT(_VaeEncoder(), spec)
    assert q2 is None and torch.equal(lat2, lat) and np.array_equal(imgs2[0], imgs[0]) and np.array_equal(signals2[0], signals[0])


def test_griffinlim_matches_oracle_and_converges():
    """Griffin-Lim from the fused STFT kernels vs the numpy restatement of librosa's loop, same initial phases (fp32 vs
    float64: 2e-3 on the waveform after 8 iterations), and the property the algorithm exists for: the magnitude of the
    result's STFT approaches the target."""
    from diffusynth_b200 import codec
    T = 64
    ys = [cases.synthetic_wave(n=256 * (T - 1), seed=s) for s in (34, 35)]
    S = np.stack([np.abs(O.stft(y)) for y in ys])
    S[:, 0] = 0.0                                   # the reference's helpers leave the DC row zero (tools.py:205-210)
    ph = 2 * np.pi * np.random.default_rng(7).random(S.shape)
    mag = torch.from_numpy(S[:, 1:]).float().cuda()
    wave = codec.griffinlim(mag, n_iter=8, init_phase=torch.from_numpy(ph[:, 1:]).float().cuda()).cpu()
    assert tuple(wave.shape) == (2, 256 * (T - 1))
    for b in range(2):
        ref = O.griffinlim(S[b], ph[b], n_iter=8)
        e = rel(wave[b], torch.from_numpy(ref))
        print(f"\ngriffinlim[{b}] waveform rel-L2 vs oracle {e:.2e}")
        assert e < 2e-3

    def spectral_error(w, b):
        return np.linalg.norm(np.abs(O.stft(w.double().numpy()))[1:] - S[b, 1:]) / np.linalg.norm(S[b, 1:])

    w0 = codec.griffinlim(mag, n_iter=0, init_phase=torch.from_numpy(ph[:, 1:]).float().cuda()).cpu()
    w32 = codec.griffinlim(mag, n_iter=32, init_phase=torch.from_numpy(ph[:, 1:]).float().cuda()).cpu()
    e0, e8, e32 = spectral_error(w0[0], 0), spectral_error(wave[0], 0), spectral_error(w32[0], 0)
    print(f"spectral convergence: {e0:.3f} -> {e8:.3f} -> {e32:.3f}")
    assert e32 < e8 < e0 and e32 < 0.25 * e0
    rnd = codec.griffinlim(mag, n_iter=4, generator=torch.Generator(device="cuda").manual_seed(1))     # random start like librosa's default
    assert torch.isfinite(rnd).all()


def test_numpy_dropins_of_tools_stft_helpers(golden):
    """codec.decode_stft / depad_STFT / encode_stft / pad_STFT (numpy in, numpy out) against the reference's own outputs
    (tests/golden/codec.npz, minted from tools.py:170-191,320-345), and codec.istft against the float64 oracle."""
    from diffusynth_b200 import codec
    g = golden["codec"]
    D = codec.decode_stft(g["decode_in"])
    assert D.dtype == g["decode_out"].dtype and D.shape == g["decode_out"].shape
    assert np.abs(D - g["decode_out"]).max() <= 2e-6 * np.abs(g["decode_out"]).max()
    dp = codec.depad_STFT(g["decode_out"])
    assert dp.dtype == g["depad_out"].dtype and np.array_equal(dp, g["depad_out"])
    pad = codec.pad_STFT(g["encode_in"], 8)
    assert np.array_equal(pad, g["pad_out"]) and codec.pad_STFT(g["encode_in"], None).shape == (512, 5) and codec.pad_STFT(g["encode_in"], 3).shape == (512, 5)
    enc = codec.encode_stft(g["pad_out"])
    assert enc.dtype == g["encode_out"].dtype and enc.shape == g["encode_out"].shape
    assert np.abs(enc - g["encode_out"]).max() < 1e-12 if enc.dtype == np.float64 else np.abs(enc - g["encode_out"]).max() < 2e-6
    Dc = (cases.randn((513, 12), 44) + 1j * cases.randn((513, 12), 45)).numpy()
    Dc[0] = 0
    w = codec.istft(Dc)
    ref = O.istft(Dc)
    assert w.shape == ref.shape and rel(torch.from_numpy(w), torch.from_numpy(ref)) < 1e-5
    a = codec.adjust_audio_length(np.ones(10), 16)
    assert a.dtype == np.float64 and a.shape == (16,) and a[:10].sum() == 10 and a[10:].sum() == 0
    assert codec.adjust_audio_length(np.ones(32000), 16000, 32000, 16000).shape == (16000,)

"""Parity AT THE BENCHMARK'S OWN CONFIGURATION: deployed U-Net, 20 respaced steps, CFG 6 -- DDIM and DDPM -- and BASELINE
config 5 (timbre modification: 20 steps / strength 0.7 -> 28 respaced steps, 19 U-Net steps, DiffSynthSampler.py:458-476,
sound2sound_with_text.py:185).  The CUDA graph loop is compared per step with the CPU oracle on the same host noise
(free-running), every step is repeated teacher-forced from the oracle's x_t (eps_u, eps_c, x_{t-1}), the results are checked
against tests/golden/headline.npz minted from the UNMODIFIED reference, and the text-to-timbre tail (VQ -> decoder -> iSTFT) is
checked on the 20-step latents.  With random-init weights the latent rms grows to ~216 (DDIM) / ~361 (DDPM) by step 20
(absmax ~1000-1500): the range checks below look at the fp16 activations of exactly those steps.
Tolerances: 1e-2 relative L2 for everything through 16-bit MMA (north_star), 1e-5 for the fp32 update kernel."""
import numpy as np
import pytest
import torch

from diffusynth_b200 import ops, weights as W
from oracle import cases, ds_oracle as O
from tests.gpu_util import rel

pytestmark = pytest.mark.gpu
TOL = 1e-2
FP16_MAX = 65504.0


@pytest.fixture(scope="module")
def pipe():
    from diffusynth_b200 import TextToTimbre
    return TextToTimbre.random_init(device="cuda", seed=0, perturb_norm=True)


@pytest.fixture(scope="module")
def usd():
    return W.unet_random_state_dict(seed=0)


def _oracle_run(name, usd, I, trace):
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, I["n_steps"], dtype=np.int32)))
    with torch.no_grad():
        ref = O.sample_loop(lambda x, t, c: O.unet_forward(usd, x, t, c), sch, (I["B"], 4, 128, 64), I["cond"], I["uncond"], 6, I["draws"],
                            sampler=I["sampler"], guide=I["guide"], start_ratio=I["strength"], trace=trace)
    return sch, ref


@pytest.mark.parametrize("name", ["ddim20", "ddpm20", "modify20"])
def test_headline_config_parity(name, pipe, usd, golden):
    from diffusynth_b200 import DiffSynthSampler
    I = cases.headline_inputs(name)
    B = I["B"]
    g = golden["headline"]
    net = pipe.unet
    s = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=B)
    s.activate_classifier_free_guidance(6, I["uncond"].cuda())
    s.respace(list(np.linspace(0, 999, I["n_steps"], dtype=np.int32)))
    s.noise_feed = I["draws"][1:]
    shape = (B, 4, 128, 64)
    if name == "modify20":
        imgs, _ = s.img_guided_sample(net, shape, I["strength"], I["guide"].cuda(), return_tensor=True, condition=I["cond"].cuda(),
                                      sampler=I["sampler"], initial_noise=I["draws"][0].cuda())
    else:
        imgs, _ = s.sample(net, shape, return_tensor=True, condition=I["cond"].cuda(), sampler=I["sampler"], initial_noise=I["draws"][0].cuda())
    assert s.last_graph_launches > 0, "the CUDA-graph loop did not run"
    n_unet = int(I["n_steps"] * I["strength"])
    assert len(imgs) == n_unet + 1 == g[f"{name}_sub"].shape[0]
    assert all(bool(torch.isfinite(im).all()) for im in imgs)

    # ---- the reference's own run (golden): every step (subsampled) and the final latent ----
    sub = np.stack([im.flatten()[::64].cpu().numpy() for im in imgs])
    e_gold_steps = [float(np.linalg.norm(sub[k] - g[f"{name}_sub"][k]) / np.linalg.norm(g[f"{name}_sub"][k])) for k in range(len(imgs))]
    e_gold = rel(imgs[-1], torch.from_numpy(g[f"{name}_final"]))
    print(f"\n[{name}] vs reference golden: final latent rel-L2 {e_gold:.2e}; per step {['%.1e' % e for e in e_gold_steps]}")

    # ---- oracle: free-running per step ----
    trace = []
    sch, ref = _oracle_run(name, usd, I, trace)
    assert len(ref) == len(imgs) and len(trace) == n_unet
    assert rel(ref[-1], torch.from_numpy(g[f"{name}_final"])) < 1e-4, "oracle drifted from the reference golden"
    errs = [rel(a, b) for a, b in zip(imgs, ref)]
    print(f"[{name}] free-running latent rel-L2 per step: {['%.1e' % e for e in errs]}")
    print(f"[{name}] latent rms per step: {['%.1f' % float(im.pow(2).mean().sqrt()) for im in imgs]}")

    # ---- teacher-forced: every step from the oracle's x_t ----
    u = I["uncond"].unsqueeze(0).repeat(B, 1)
    eta = 0.0 if I["sampler"] == "ddim" else 1.0
    worst = dict(eps_u=0.0, eps_c=0.0, x_prev=0.0, absmax=0.0, saturated=0)
    for k, tr in enumerate(trace):
        t_mapped = torch.full((2 * B,), sch.timestep_map[tr["t"]], dtype=torch.long)
        taps = {}
        eps = net.forward(torch.cat([tr["x"], tr["x"]]).cuda(), t_mapped.cuda(), torch.cat([u, I["cond"]]).cuda(), taps=taps)
        e_u, e_c = rel(eps[:B], tr["eps_u"]), rel(eps[B:], tr["eps_c"])
        z = I["draws"][1 + k].cuda().contiguous()
        coef = torch.tensor(s._coef(tr["t"], eta), dtype=torch.float32, device="cuda")
        nxt = torch.empty((B, 4, 128, 64), dtype=torch.float32, device="cuda")
        ops.ddim_step(eps[:B].contiguous(), eps[B:].contiguous(), tr["x"].cuda().contiguous(), z, coef, nxt)
        e_x = rel(nxt, tr["x_prev"])
        amax = max(float(v.abs().max()) for name_, v in taps.items() if v.dim() == 4)
        sat = sum(int((v.abs() >= FP16_MAX).sum()) for v in taps.values() if v.dim() == 4)
        for key, val in (("eps_u", e_u), ("eps_c", e_c), ("x_prev", e_x), ("absmax", amax)):
            worst[key] = max(worst[key], val)
        worst["saturated"] += sat
        print(f"    step {k:2d} (t={tr['t']:2d}): teacher-forced eps_u {e_u:.2e} eps_c {e_c:.2e} x_prev {e_x:.2e}; max|activation| {amax:.1f}, saturated {sat}")
        assert e_u < TOL and e_c < TOL and e_x < TOL, (k, e_u, e_c, e_x)
    print(f"[{name}] worst teacher-forced: {worst}")
    assert worst["saturated"] == 0 and worst["absmax"] < FP16_MAX / 4, worst
    assert max(errs) < TOL, errs
    assert e_gold < TOL and max(e_gold_steps) < TOL


def test_headline_tail_on_20_step_latents(pipe, usd):
    """The text-to-timbre tail at the headline step count: TextToTimbre.generate (20 DDIM steps, CFG 6) -> quantiser ->
    decoder -> iSTFT against the oracle tail; VQ indices bit-exact given the CUDA latents."""
    I = cases.headline_inputs("ddim20")
    vsd = W.vqgan_random_state_dict(seed=1)
    _, dec_plan = W.vqgan_layer_plan(W.VQGAN_DEPLOYED)
    out = pipe.generate(I["cond"].cuda(), I["uncond"].cuda(), steps=20, cfg_scale=6, noise_feed=I["draws"])
    with torch.no_grad():
        q_same, idx_same = O.vq_quantize(out.latents.cpu(), vsd["_vq_vae._embedding.weight"])
        spec_ref = O.vqgan_decode(vsd, dec_plan, q_same)
    wave_ref = np.stack([O.spectrogram_to_waveform(sp.numpy().astype(np.float64)) for sp in spec_ref])
    assert torch.equal(q_same, out.quantized.cpu())
    e_spec, e_wave = rel(out.spectrograms, spec_ref), rel(out.waveforms, torch.from_numpy(wave_ref))
    print(f"\n20-step tail: spectrogram {e_spec:.2e}  waveform {e_wave:.2e}")
    assert bool(torch.isfinite(out.waveforms).all()) and tuple(out.waveforms.shape) == (2, 65280)
    assert e_spec < TOL and e_wave < 2 * TOL


@pytest.mark.parametrize("scale", [1.0, 8.0, 64.0])
def test_unet_input_range_fp16(scale, pipe, usd):
    """What the fp16 activation format does with large latents: the final-step input of the 20-step DDPM run (rms ~361, the
    largest the benchmark produces) scaled by 1, 8 and 64.  GroupNorm(1,C) sits after every depthwise conv and in front of every
    attention, so only the residual stream scales with the input; fp16 (max 65504, packs saturate instead of overflowing to inf)
    holds x8 with the same accuracy; x64 (absmax ~1e5 at the input) is outside the representable range of the first activation
    tensor: the output must stay finite and the error is reported."""
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "headline.npz"))
    x = torch.from_numpy(g["ddpm20_final"]) * scale
    B = x.shape[0]
    cond, uncond = W.synthetic_conditions(B, 512)
    t = torch.full((B,), 0, dtype=torch.long)
    taps = {}
    eps = pipe.unet.forward(x.cuda(), t.cuda(), cond.cuda(), taps=taps).cpu()
    with torch.no_grad():
        ref = O.unet_forward(usd, x, t, cond)
    amax = max(float(v.abs().max()) for v in taps.values() if v.dim() == 4)
    sat = sum(int((v.abs() >= FP16_MAX).sum()) for v in taps.values() if v.dim() == 4)
    e = rel(eps, ref)
    print(f"\n[input x{scale:g}] input absmax {float(x.abs().max()):.0f}, max|activation| {amax:.0f}, saturated values {sat}, eps rel-L2 {e:.2e}")
    assert bool(torch.isfinite(eps).all())
    if scale <= 8.0:
        assert sat == 0 and e < TOL

"""Model-level parity on the GPU: U-Net forward, VQGAN decoder/encoder, the sampling loop (teacher-forced and
free-running) and the full text-to-timbre pipeline, CUDA path (through the C ABI) vs the CPU oracle.
Tolerance for everything that passes through bf16 tensor-core math: relative L2 <= 1e-2 against the fp32
oracle (BASELINE.json north_star); fp32 kernels: <= 1e-5; codebook indices: bit-exact."""
import numpy as np
import pytest
import torch

from diffusynth_b200 import weights as W
from oracle import cases, ds_oracle as O
from tests.gpu_util import rel

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2


def _unet(cfg, sd):
    from diffusynth_b200 import ConditionedUnet
    net = ConditionedUnet(**{k: v for k, v in cfg.items() if k not in ("out_dim", "time_dim")}, device="cuda")
    net.load_state_dict(sd)
    return net


@pytest.mark.parametrize("name", ["small_w10", "deployed_w28", "small_cat_w16", "small_resnet_w16"])
def test_unet_odd_width_parity(name, golden):
    """Widths whose stride-2 levels are odd (track_maker's per-note widths, track_maker.py:245): pad_to_match; and the
    attn_type="linear_cat" variant (the condition as an extra key / value token, ds_attn_finalize_cat) and the ResnetBlock variant
    (use_convnext=False: conv3x3 -> GroupNorm(8) -> SiLU blocks)."""
    cfg, sd, x, t, cond = cases.unet_case(name)
    net = _unet(cfg, sd)
    eps = net.forward(x.cuda(), t.cuda(), cond.cuda()).cpu()
    with torch.no_grad():
        ref = O.unet_forward(sd, x, t, cond)
    print(f"\n[{name}] eps rel-L2 {rel(eps, ref):.3e}")
    assert rel(eps, ref) < BF16_TOL
    assert rel(eps, torch.from_numpy(golden["extra"][f"{name}_eps"])) < BF16_TOL       # the reference's own output


@pytest.mark.parametrize("name", ["small_w16", "deployed_w64", "deployed_w24"])
def test_unet_forward_parity(name, golden):
    cfg, sd, x, t, cond = cases.unet_case(name)
    net = _unet(cfg, sd)
    taps, ref_taps = {}, {}
    eps = net.forward(x.cuda(), t.cuda(), cond.cuda(), taps=taps).cpu()
    with torch.no_grad():
        ref = O.unet_forward(sd, x, t, cond, ref_taps)
    report = [(k, rel(taps[k], ref_taps[k])) for k in ref_taps if k in taps]
    worst = max(report, key=lambda kv: kv[1])
    print(f"\n[{name}] eps rel-L2 {rel(eps, ref):.3e}; worst layer {worst[0]} {worst[1]:.3e}")
    for k, v in report:
        print(f"    {k:16s} {v:.3e}")
    assert rel(eps, ref) < BF16_TOL
    assert all(v < 2 * BF16_TOL for _, v in report), worst
    assert rel(eps, torch.from_numpy(golden["unet"][f"{name}_eps"])) < BF16_TOL       # the reference's own output


@pytest.mark.parametrize("name", ["small_family_w16", "small_notime_w16", "small_nocond_w16", "small_cat_nocond_w16"])
def test_unet_conditioning_variants_parity(name, golden):
    """condition_type="instrument_family" (integer labels, ds_embedding_gather), with_time_emb=False and condition=None
    (diffusion_components.py:155-168; diffusion.py:107-109,199-202,211)."""
    cfg, sd, x, t, cond = cases.unet_case(name)
    net = _unet(cfg, sd)
    eps = net.forward(x.cuda(), t.cuda(), None if cond is None else cond.cuda()).cpu()
    with torch.no_grad():
        ref = O.unet_forward(sd, x, t, cond)
    print(f"\n[{name}] eps rel-L2 {rel(eps, ref):.3e}")
    assert rel(eps, ref) < BF16_TOL
    assert rel(eps, torch.from_numpy(golden["variants"][f"{name}_eps"])) < BF16_TOL       # the reference's own output


def test_unet_rejects_unsupported_variants():
    from diffusynth_b200 import ConditionedUnet
    with pytest.raises(NotImplementedError):
        ConditionedUnet(in_dim=4, attn_type="bogus")
    with pytest.raises(NotImplementedError):
        ConditionedUnet(in_dim=4, attn_type="linear_add", condition_type="bogus")


def test_vqgan_decoder_encoder_parity(golden):
    from diffusynth_b200 import VQGAN
    sd = W.vqgan_random_state_dict(seed=1)
    vq = VQGAN(**W.VQGAN_DEPLOYED, device="cuda")
    vq.load_state_dict(sd)
    enc_plan, dec_plan = W.vqgan_layer_plan(W.VQGAN_DEPLOYED)
    lat = cases.vq_latents(B=2)
    q, loss, (perp, _, _) = vq._vq_vae(lat.cuda())
    q_ref, _ = O.vq_quantize(lat, sd["_vq_vae._embedding.weight"])
    assert torch.equal(q.cpu(), q_ref) and torch.isfinite(loss) and torch.isfinite(perp)
    dec = vq._decoder(q).cpu()
    with torch.no_grad():
        dec_ref = O.vqgan_decode(sd, dec_plan, q_ref)
    print(f"\ndecoder rel-L2 {rel(dec, dec_ref):.3e} (mag {rel(dec[:, 0], dec_ref[:, 0]):.3e}, cos {rel(dec[:, 1], dec_ref[:, 1]):.3e})")
    assert tuple(dec.shape) == (2, 3, 512, 256) and rel(dec, dec_ref) < BF16_TOL
    assert rel(dec[:1].flatten()[::7], torch.from_numpy(golden["vqgan"]["dec_sub"])) < BF16_TOL
    spec = torch.from_numpy(O.waveform_to_spectrogram(cases.synthetic_wave())[None])
    enc = vq._encoder(spec.cuda()).cpu()
    with torch.no_grad():
        enc_ref = O.vqgan_encode(sd, enc_plan, spec)
    print(f"encoder rel-L2 {rel(enc, enc_ref):.3e}")
    assert tuple(enc.shape) == (1, 4, 128, 64) and rel(enc, enc_ref) < BF16_TOL
    assert rel(enc, torch.from_numpy(golden["vqgan"]["enc_lat"])) < BF16_TOL


@pytest.mark.parametrize("kind", ["ddim", "ddpm", "nocfg", "guided", "inpaint"])
def test_sampler_generic_model_matches_reference_golden(kind, golden):
    """Drop-in behaviour with an arbitrary callable model: same loops, same outputs as the reference (fp32 kernels)."""
    from diffusynth_b200 import DiffSynthSampler
    g = golden["sampler"]
    B, Wd = 3, 40
    draws = cases.randn((12, B, 4, 128, 64), 6)
    cond, uncond = W.synthetic_conditions(B, 16, seed=77)
    guide = cases.randn((B, 4, 128, 64), 8) * 0.5
    s = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=B)
    s.noise_feed = draws[1:]
    model = lambda x, t, c: cases.toy_model(x, t, c)
    if kind != "nocfg":
        s.activate_classifier_free_guidance(6, uncond.cuda())
    if kind == "guided":
        s.respace(list(np.linspace(0, 999, int(8 / 0.7), dtype=np.int32)))
        imgs, _ = s.img_guided_sample(model, (B, 4, 128, 64), 0.7, guide.cuda(), return_tensor=True, condition=cond.cuda(), initial_noise=draws[0].cuda())
        assert len(imgs) == int(g["loop_guided_len"]) and rel(imgs[0], torch.from_numpy(g["loop_guided_first"])) < 1e-6
    else:
        s.respace(list(np.linspace(0, 999, 8, dtype=np.int32)))
        if kind == "inpaint":
            mask = (cases.randn((B, 1, 128, Wd), 9) > 0).float()
            imgs, _ = s.inpaint_sample(model, (B, 4, 128, Wd), 1.0, guide.cuda(), mask.cuda(), return_tensor=True, condition=cond.cuda(), initial_noise=draws[0].cuda())
        else:
            imgs, _ = s.sample(model, (B, 4, 128, Wd), return_tensor=True, condition=cond.cuda(), initial_noise=draws[0].cuda(),
                               sampler="ddpm" if kind == "ddpm" else "ddim")
            assert len(imgs) == 9
    assert rel(imgs[-1], torch.from_numpy(g[f"loop_{kind}_last"])) < 1e-5
    with pytest.raises(NotImplementedError):
        s.sample(model, (B, 4, 128, Wd), condition=cond.cuda(), sampler="euler")


@pytest.mark.parametrize("sampler_kind", ["ddim", "ddpm"])
def test_graph_sampling_loop_parity(sampler_kind):
    """CUDA-graph loop with the B200 U-Net vs the oracle loop on the same host noise: free-running latents per
    step, and teacher-forced eps / x_{t-1} (feed the oracle's x_t into one step)."""
    from diffusynth_b200 import DiffSynthSampler
    cfg, sd, _, _, _ = cases.unet_case("deployed_w64")
    net = _unet(cfg, sd)
    B, steps = 2, 4
    draws = W.host_noise(3, 1 + steps, B)
    cond, uncond = W.synthetic_conditions(B, 512)
    s = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=B)
    s.activate_classifier_free_guidance(6, uncond.cuda())
    s.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
    s.noise_feed = draws[1:]
    imgs, init = s.sample(net, (B, 4, 128, 64), return_tensor=True, condition=cond.cuda(), sampler=sampler_kind, initial_noise=draws[0].cuda())
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
    trace = []
    with torch.no_grad():
        ref = O.sample_loop(lambda x, t, c: O.unet_forward(sd, x, t, c), sch, (B, 4, 128, 64), cond, uncond, 6, draws, sampler=sampler_kind, trace=trace)
    assert len(imgs) == len(ref) == steps + 1 and torch.equal(imgs[0].cpu(), ref[0])
    errs = [rel(a, b) for a, b in zip(imgs, ref)]
    print(f"\n[{sampler_kind}] free-running latent rel-L2 per step: {['%.2e' % e for e in errs]}")
    # teacher-forced: one step from the oracle's x_t
    for k, tr in enumerate(trace):
        t_mapped = torch.full((2 * B,), sch.timestep_map[tr["t"]], dtype=torch.long)
        u = uncond.unsqueeze(0).repeat(B, 1)
        eps = net(torch.cat([tr["x"], tr["x"]]).cuda(), t_mapped.cuda(), torch.cat([u, cond]).cuda()).cpu()
        e_u, e_c = rel(eps[:B], tr["eps_u"]), rel(eps[B:], tr["eps_c"])
        print(f"    step {k} (t={tr['t']}): teacher-forced eps_u {e_u:.2e} eps_c {e_c:.2e}")
        assert e_u < BF16_TOL and e_c < BF16_TOL
    assert max(errs) < BF16_TOL
    # replaying the cached graph with the same inputs is deterministic
    s.noise_feed = draws[1:]
    imgs2, _ = s.sample(net, (B, 4, 128, 64), return_tensor=True, condition=cond.cuda(), sampler=sampler_kind, initial_noise=draws[0].cuda())
    assert all(torch.equal(a, b) for a, b in zip(imgs, imgs2))


def test_text_to_timbre_pipeline_parity():
    """sampling -> VQ -> decoder -> iSTFT at B=2, 3 steps: spectrogram and waveform vs the oracle pipeline; VQ indices
    bit-exact given identical quantiser inputs."""
    from diffusynth_b200 import TextToTimbre
    from diffusynth_b200 import ConditionedUnet, VQGAN
    usd, vsd = W.unet_random_state_dict(seed=0), W.vqgan_random_state_dict(seed=1)
    unet = _unet(W.UNET_DEPLOYED, usd)
    vq = VQGAN(**W.VQGAN_DEPLOYED, device="cuda")
    vq.load_state_dict(vsd)
    pipe = TextToTimbre(unet, vq)
    B, steps = 2, 3
    draws = W.host_noise(5, 1 + steps, B)
    cond, uncond = W.synthetic_conditions(B, 512)
    out = pipe.generate(cond.cuda(), uncond.cuda(), steps=steps, cfg_scale=6, noise_feed=draws)
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
    _, dec_plan = W.vqgan_layer_plan(W.VQGAN_DEPLOYED)
    with torch.no_grad():
        lat_ref = O.sample_loop(lambda x, t, c: O.unet_forward(usd, x, t, c), sch, (B, 4, 128, 64), cond, uncond, 6, draws)[-1]
        # quantiser: identical inputs (the CUDA latents) -> identical indices
        q_same, idx_same = O.vq_quantize(out.latents.cpu(), vsd["_vq_vae._embedding.weight"])
        assert torch.equal(q_same, out.quantized.cpu()) and torch.equal(idx_same, pipe.vqgan._vq_vae.last_indices.cpu())
        spec_ref = O.vqgan_decode(vsd, dec_plan, q_same)
    wave_ref = np.stack([O.spectrogram_to_waveform(s.numpy().astype(np.float64)) for s in spec_ref])
    e_lat, e_spec, e_wave = rel(out.latents, lat_ref), rel(out.spectrograms, spec_ref), rel(out.waveforms, torch.from_numpy(wave_ref))
    # the fp32 iSTFT alone, on the CUDA spectrogram
    wave_own = np.stack([O.spectrogram_to_waveform(s.numpy().astype(np.float64)) for s in out.spectrograms.cpu()])
    e_istft = rel(out.waveforms, torch.from_numpy(wave_own))
    print(f"\npipeline: latent {e_lat:.2e}  spectrogram {e_spec:.2e}  waveform {e_wave:.2e}  (iSTFT alone {e_istft:.2e})")
    assert tuple(out.waveforms.shape) == (B, 65280)
    assert e_lat < BF16_TOL and e_spec < BF16_TOL and e_wave < 2 * BF16_TOL and e_istft < 1e-5


def test_timbre_modification_pipeline_parity():
    """BASELINE config 5 at B=2: waveform -> STFT+ -> encoder -> q_sample(guide, strength) -> partial re-denoise with CFG ->
    VQ -> decoder -> iSTFT.  strength 0.7, 3 steps -> respaced to int(3/0.7) = 4, start index int(4*0.7) = 2 -> 2 U-Net steps."""
    from diffusynth_b200 import TextToTimbre, VQGAN
    usd, vsd = W.unet_random_state_dict(seed=0), W.vqgan_random_state_dict(seed=1)
    unet = _unet(W.UNET_DEPLOYED, usd)
    vq = VQGAN(**W.VQGAN_DEPLOYED, device="cuda")
    vq.load_state_dict(vsd)
    pipe = TextToTimbre(unet, vq)
    B, steps, strength = 2, 3, 0.7
    wave = torch.from_numpy(cases.synthetic_wave(seed=33)).float()[None]
    guide = pipe.encode_audio(wave.cuda())
    enc_plan, dec_plan = W.vqgan_layer_plan(W.VQGAN_DEPLOYED)
    with torch.no_grad():
        spec_ref = torch.from_numpy(O.waveform_to_spectrogram(cases.synthetic_wave(seed=33) / np.abs(cases.synthetic_wave(seed=33)).max())[None])
        guide_ref = O.vqgan_encode(vsd, enc_plan, spec_ref)
    e_guide = rel(guide, guide_ref)
    n_steps = int(steps / strength)
    draws = W.host_noise(9, 1 + n_steps, B)
    cond, uncond = W.synthetic_conditions(B, 512)
    out = pipe.modify(guide_ref.cuda(), cond.cuda(), uncond.cuda(), steps=steps, strength=strength, noise_feed=draws)
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, n_steps, dtype=np.int32)))
    with torch.no_grad():
        ref = O.sample_loop(lambda x, t, c: O.unet_forward(usd, x, t, c), sch, (B, 4, 128, 64), cond, uncond, 6, draws,
                            guide=guide_ref.repeat(B, 1, 1, 1), start_ratio=strength)
        q_same, _ = O.vq_quantize(out.latents.cpu(), vsd["_vq_vae._embedding.weight"])
        spec = O.vqgan_decode(vsd, dec_plan, q_same)
    wave_ref = np.stack([O.spectrogram_to_waveform(s.numpy().astype(np.float64)) for s in spec])
    e_lat, e_spec, e_wave = rel(out.latents, ref[-1]), rel(out.spectrograms, spec), rel(out.waveforms, torch.from_numpy(wave_ref))
    print(f"\nmodification: guide latent {e_guide:.2e}  ({len(ref) - 1} U-Net steps) latent {e_lat:.2e}  spectrogram {e_spec:.2e}  waveform {e_wave:.2e}")
    assert len(ref) == int(n_steps * strength) + 1
    assert torch.equal(q_same, out.quantized.cpu())
    assert e_guide < BF16_TOL and e_lat < BF16_TOL and e_spec < BF16_TOL and e_wave < 2 * BF16_TOL


def test_graph_loop_class_label_conditioning():
    """condition_type="instrument_family" through the library's sampling graph with classifier-free guidance: integer labels,
    the unconditional condition is a scalar label (DiffSynthSampler.py:313-317 repeats it per sample)."""
    from diffusynth_b200 import DiffSynthSampler
    cfg, sd, _, _, _ = cases.unet_case("small_family_w16")
    net = _unet(cfg, sd)
    B, steps, Hh, Wd = 3, 4, 32, 32
    draws = cases.randn((1 + steps, B, 4, Hh, 64), 64)
    cond, uncond = torch.tensor([3, 0, 9]), torch.tensor(11)
    s = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=B, height=Hh)
    s.noise_feed = draws[1:]
    s.activate_classifier_free_guidance(4, uncond.cuda())
    s.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
    imgs, _ = s.sample(net, (B, 4, Hh, Wd), return_tensor=True, condition=cond.cuda(), initial_noise=draws[0].cuda())
    assert s.last_graph_launches > 0 and next(iter(s._graphs.values())).sgraph is not None
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
    with torch.no_grad():
        ref = O.sample_loop(lambda x, t, c: O.unet_forward(sd, x, t, c), sch, (B, 4, Hh, Wd), cond, uncond, 4, draws)
    errs = [rel(a, b) for a, b in zip(imgs, ref)]
    print(f"\n[class labels] latents rel-L2 per step: {[f'{e:.1e}' for e in errs]}")
    assert max(errs) < BF16_TOL


@pytest.mark.parametrize("mode", ["inpaint", "guided", "nocfg", "wide"])
def test_graph_loop_variants_small_unet(mode):
    """The CUDA-graph loop with the B200 U-Net in its other modes (inpaint blend with a fixed mask, image-guided start,
    CFG off, latent width != 64) against the oracle loop; small U-Net so the CPU oracle stays fast."""
    from diffusynth_b200 import DiffSynthSampler
    cfg, sd, _, _, _ = cases.unet_case("small_w16")
    net = _unet(cfg, sd)
    B, steps, Hh = 3, 5, 32
    Wd = 40 if mode == "wide" else (64 if mode == "guided" else 32)
    draws = cases.randn((1 + steps + 2, B, 4, Hh, 64), 61)
    cond, uncond = W.synthetic_conditions(B, 64, seed=300)
    guide = cases.randn((B, 4, Hh, 64), 62) * 0.5
    s = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=B, height=Hh)
    sch = O.Schedule(1000)
    model = lambda x, t, c: O.unet_forward(sd, x, t, c)
    s.noise_feed = draws[1:]
    if mode != "nocfg":
        s.activate_classifier_free_guidance(4, uncond.cuda())
    cfg_scale, unc = (1.0, None) if mode == "nocfg" else (4, uncond)
    if mode == "guided":
        n = int(steps / 0.6)
        s.respace(list(np.linspace(0, 999, n, dtype=np.int32)))
        sch.respace(list(np.linspace(0, 999, n, dtype=np.int32)))
        imgs, _ = s.img_guided_sample(net, (B, 4, Hh, Wd), 0.6, guide.cuda(), return_tensor=True, condition=cond.cuda(),
                                      initial_noise=draws[0].cuda(), sampler="ddpm")
        with torch.no_grad():
            ref = O.sample_loop(model, sch, (B, 4, Hh, Wd), cond, unc, cfg_scale, draws, guide=guide, start_ratio=0.6, sampler="ddpm")
    else:
        s.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
        sch.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
        if mode == "inpaint":
            mask = (cases.randn((B, 1, Hh, Wd), 63) > 0).float()
            imgs, _ = s.inpaint_sample(net, (B, 4, Hh, Wd), 1.0, guide.cuda(), mask.cuda(), return_tensor=True, condition=cond.cuda(),
                                       initial_noise=draws[0].cuda())
            with torch.no_grad():
                ref = O.sample_loop(model, sch, (B, 4, Hh, Wd), cond, unc, cfg_scale, draws, guide=guide, mask=mask, inpaint=True)
        else:
            imgs, _ = s.sample(net, (B, 4, Hh, Wd), return_tensor=True, condition=cond.cuda(), initial_noise=draws[0].cuda())
            with torch.no_grad():
                ref = O.sample_loop(model, sch, (B, 4, Hh, Wd), cond, unc, cfg_scale, draws)
    assert len(imgs) == len(ref)
    errs = [rel(a, b) for a, b in zip(imgs, ref)]
    print(f"\n[{mode}] W={Wd} latents rel-L2 per step: {['%.1e' % e for e in errs]}")
    assert max(errs) < BF16_TOL


def test_dynamic_mask_inpaint_generic_model_matches_reference_golden(golden):
    """inpaint_sample(use_dynamic_mask=True) with an arbitrary model vs the reference's own output (fp32 kernels)."""
    from diffusynth_b200 import DiffSynthSampler
    g = golden["extra"]
    B, Wd, Hh = 2, 150, 16
    draws = cases.randn((10, B, 4, Hh, 64), 71)
    cond, uncond = W.synthetic_conditions(B, 16, seed=78)
    guide = cases.randn((B, 4, Hh, 64), 72) * 0.5
    s = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=B, height=Hh)
    s.noise_feed = draws[1:]
    s.activate_classifier_free_guidance(3, uncond.cuda())
    s.respace(list(np.linspace(0, 999, 8, dtype=np.int32)))
    imgs, _ = s.inpaint_sample(lambda x, t, c: cases.toy_model(x, t, c), (B, 4, Hh, Wd), 1.0, guide.cuda(), None, return_tensor=True,
                               condition=cond.cuda(), initial_noise=draws[0].cuda(), use_dynamic_mask=True, mask_flexivity=0.8)
    assert rel(imgs[-1], torch.from_numpy(g["loop_dynmask_last"])) < 1e-5 and rel(imgs[4], torch.from_numpy(g["loop_dynmask_mid"])) < 1e-5
    _, pts = s.get_deterministic_noise_tensor_repeat(B, Wd, reference_noise=guide.cuda())
    for k, m in enumerate(s.get_dynamic_masks(8, (B, 4, Hh, Wd), pts, 0.8)):
        assert np.array_equal(m[0, 0, 0].cpu().numpy().astype(np.uint8), g[f"dynmask_{k}"])


def test_note_synthesis_odd_width_parity():
    """track_maker's per-note synthesis (track_maker.py:228-283) through TextToTimbre.synthesize_note: duration 0.9 s -> width 30
    (levels 30/15/7/3, two of them odd), no guidance, dynamic-mask inpainting from the instrument's latent, then VQ -> decoder ->
    iSTFT, against the oracle loop on the same host noise."""
    from diffusynth_b200 import TextToTimbre, VQGAN
    usd, vsd = W.unet_random_state_dict(seed=0), W.vqgan_random_state_dict(seed=1)
    vq = VQGAN(**W.VQGAN_DEPLOYED, device="cuda")
    vq.load_state_dict(vsd)
    pipe = TextToTimbre(_unet(W.UNET_DEPLOYED, usd), vq)
    _, dec_plan = W.vqgan_layer_plan(W.VQGAN_DEPLOYED)
    steps, dur = 4, 0.9
    Wd = int(256 * ((dur + 1) / 4) / 4)
    assert Wd == 30
    draws = W.host_noise(21, 1 + steps, 1)
    cond, _ = W.synthetic_conditions(1, 512)
    inst = cases.vq_latents(B=1, seed=22)
    out = pipe.synthesize_note(inst.cuda(), cond.cuda(), dur, sample_steps=steps, noise_feed=draws)
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
    with torch.no_grad():
        ref = O.sample_loop(lambda x, t, c: O.unet_forward(usd, x, t, c), sch, (1, 4, 128, Wd), cond, None, 1.0, draws, guide=inst,
                            inpaint=True, dynamic_mask_flexivity=1.0)
        q_same, _ = O.vq_quantize(out.latents.cpu(), vsd["_vq_vae._embedding.weight"])
        spec = O.vqgan_decode(vsd, dec_plan, q_same)
    wave_ref = O.spectrogram_to_waveform(spec[0].numpy().astype(np.float64))
    e_lat, e_spec, e_wave = rel(out.latents, ref[-1]), rel(out.spectrograms, spec), rel(out.waveforms[0], torch.from_numpy(wave_ref))
    print(f"\nnote (W={Wd}): latent {e_lat:.2e}  spectrogram {e_spec:.2e}  waveform {e_wave:.2e}")
    assert tuple(out.spectrograms.shape) == (1, 3, 512, 4 * Wd) and tuple(out.waveforms.shape) == (1, 256 * (4 * Wd - 1))
    assert torch.equal(q_same, out.quantized.cpu())
    assert e_lat < BF16_TOL and e_spec < BF16_TOL and e_wave < 2 * BF16_TOL


def test_batched_notes_equal_single_notes():
    """notes of equal duration go through ONE batched graph launch (TextToTimbre.synthesize_notes); every note of
    the batch equals the same note synthesised on its own with the same noise (the path has no cross-sample operation).  Bit-level
    equality needs batch_invariant=True: with the default tiling policy a one-note job runs narrower N tiles than a three-note job
    (tests/test_gpu_fullsize.py compares the two policies)."""
    from diffusynth_b200 import TextToTimbre
    pipe = TextToTimbre.random_init(device="cuda", seed=0, batch_invariant=True)
    steps = 3
    durs = [0.9, 2.0, 0.9, 0.9]                       # widths 30, 48, 30, 30 -> two groups (batch 3 and batch 1)
    feeds = [W.host_noise(40 + i, 1 + steps, 1) for i in range(len(durs))]
    cond, _ = W.synthetic_conditions(1, 512)
    inst = cases.vq_latents(B=1, seed=22)
    notes = pipe.synthesize_notes(inst.cuda(), cond.cuda(), durs, sample_steps=steps, noise_feeds=feeds)
    n_graphs = sum(len(s._graphs) for k, s in pipe._samplers.items() if k[0] == "note")
    assert n_graphs == 2 and [n.waveforms.shape[1] for n in notes] == [256 * (4 * TextToTimbre.note_width(d) - 1) for d in durs]
    worst = 0.0
    for i in (0, 2, 1):
        one = pipe.synthesize_note(inst.cuda(), cond.cuda(), durs[i], sample_steps=steps, noise_feed=feeds[i])
        worst = max(worst, rel(notes[i].latents, one.latents), rel(notes[i].waveforms, one.waveforms))
        assert torch.equal(notes[i].quantized, one.quantized)
    print(f"\nbatched notes vs single notes: worst rel-L2 {worst:.2e}")
    assert worst < 1e-6
    # a second track with the same durations reuses both graphs
    pipe.synthesize_notes(inst.cuda(), cond.cuda(), durs, sample_steps=steps)
    assert sum(len(s._graphs) for k, s in pipe._samplers.items() if k[0] == "note") == 3      # (+ the single-note graph of width 30 / 48)

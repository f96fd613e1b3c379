"""Mint golden vectors from the UNMODIFIED reference (imported from /root/reference).
Run here (the reference is absent on the GPU box):  python -m oracle.make_golden
Outputs: tests/golden/*.npz.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusynth_b200 import weights as W            # noqa: E402
from oracle import cases, ref_loader                # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def stats(t):
    t = t.double()
    return np.array([t.mean().item(), t.pow(2).mean().sqrt().item()])


@torch.no_grad()
def extra(ref):
    """tests/golden/extra.npz: odd-width U-Net cases (pad_to_match) and the image products / 6-list result of the decode
    glue (utils.py:8-128,194-267), minted from the unmodified reference.  `python -m oracle.make_golden --extra` writes
    only this file."""
    g = {}
    for name in ("deployed_w28", "small_w10", "small_cat_w16", "small_resnet_w16"):
        cfg, sd, x, t, cond = cases.unet_case(name)
        net = ref.ConditionedUnet(**cfg).eval()
        net.load_state_dict(sd, strict=True)
        g[f"{name}_eps"] = net(x, t, cond).numpy()
    # inpainting with the shrinking dynamic masks (DiffSynthSampler.py:365-422,483-487) on the toy model: W = 150 has a
    # repeated middle segment, height 16 keeps the fixture small
    B, Wd, Hh = 2, 150, 16
    draws = cases.randn((10, B, 4, Hh, 64), 71)
    cond, uncond = W.synthetic_conditions(B, 16, seed=78)
    guide = cases.randn((B, 4, Hh, 64), 72) * 0.5
    S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B, height=Hh), draws)
    S.activate_classifier_free_guidance(3, uncond)
    S.respace(list(np.linspace(0, 999, 8, dtype=np.int32)))
    imgs, _ = S.inpaint_sample(cases.toy_model, (B, 4, Hh, Wd), 1.0, guide, None, return_tensor=True, condition=cond,
                               initial_noise=draws[0], use_dynamic_mask=True, mask_flexivity=0.8)
    g["loop_dynmask_last"], g["loop_dynmask_mid"] = imgs[-1].numpy(), imgs[4].numpy()
    _, pts = S.get_deterministic_noise_tensor_repeat(B, Wd, reference_noise=guide)
    for k, m in enumerate(S.get_dynamic_masks(8, (B, 4, Hh, Wd), pts, 0.8)):
        g[f"dynmask_{k}"] = m[0, 0, 0].numpy().astype(np.uint8)
    glue = ref_loader.load_glue()
    spec = cases.spec_representation()
    other = cases.spec_representation(seed=53)

    class FakeDecoder:            # the glue only calls decoder(latents) (utils.py:221-224)
        def __call__(self, z):
            return spec.clone()

    out = glue.encodeBatch2GradioOutput_STFT(FakeDecoder(), torch.zeros(2, 4, 128, 3), resolution=(512, 12),
                                             original_STFT_batch=other.numpy())
    for k, lst in zip(("mag_img", "phase_img", "signal", "mag_img_amp", "phase_img_amp", "signal_amp"), out):
        g[f"glue_{k}"] = np.stack(lst)
    lat = cases.small_latents().numpy()
    g["latent_img"] = np.stack([glue.latent_representation_to_Gradio_image(lat[b].copy()) for b in range(lat.shape[0])])
    np.savez_compressed(os.path.join(OUT, "extra.npz"), **g)


@torch.no_grad()
def headline(ref):
    """tests/golden/headline.npz: the benchmark's own sampling configuration (deployed U-Net, 20 respaced steps, CFG 6, DDIM and
    DDPM) and BASELINE config 5 (timbre modification, 20 / 0.7 -> 28 respaced, 19 U-Net steps) run through the UNMODIFIED reference
    (DiffSynthSampler.py:520-536, :562-583) at B = 2 with host-fed noise.  `python -m oracle.make_golden --headline`."""
    g = {}
    usd = W.unet_random_state_dict(seed=0)
    net = ref.ConditionedUnet(**W.UNET_DEPLOYED).eval()
    net.load_state_dict(usd, strict=True)
    for name in ("ddim20", "ddpm20", "modify20"):
        I = cases.headline_inputs(name)
        S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=I["B"]), I["draws"])
        S.activate_classifier_free_guidance(6, I["uncond"])
        S.respace(list(np.linspace(0, 999, I["n_steps"], dtype=np.int32)))
        shape = (I["B"], 4, 128, 64)
        if name == "modify20":
            imgs, _ = S.img_guided_sample(net, shape, I["strength"], I["guide"], return_tensor=True, condition=I["cond"],
                                          sampler=I["sampler"], initial_noise=I["draws"][0])
        else:
            imgs, _ = S.sample(net, shape, return_tensor=True, condition=I["cond"], sampler=I["sampler"], initial_noise=I["draws"][0])
        g[f"{name}_sub"], g[f"{name}_stats"], g[f"{name}_final"] = cases.headline_digest(imgs)
        print(name, len(imgs), "latents; rms per step", np.round(g[f"{name}_stats"][:, 1], 3))
    np.savez_compressed(os.path.join(OUT, "headline.npz"), **g)


@torch.no_grad()
def sampler2(ref):
    """tests/golden/sampler2.npz (`--sampler2`): sampler surfaces the first fixtures did not reach -- interpolate() between two noise
    endpoints (:538-560; the random-endpoint branch of generate_linear_noise raises in the reference, :247-252 unpack a 1-element
    tensor into two names), the "non_repeat" noise strategy (:62-77), q_sample with per-sample timesteps (:271-294) and inpainting
    with a per-channel [B,C,H,W] mask as inpaint_with_text.py:229-231 passes it -- all on the toy model, height 16."""
    g = {}
    B, Hh = 3, 16
    cond, uncond = W.synthetic_conditions(B, 16, seed=91)
    draws = cases.randn((12, B, 4, Hh, 64), 92)
    e0, e1 = cases.randn((4, Hh, 64), 93), cases.randn((4, Hh, 64), 94)
    S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B, height=Hh), draws)
    S.activate_classifier_free_guidance(3, uncond)
    S.respace(list(np.linspace(0, 999, 6, dtype=np.int32)))
    imgs, init = S.interpolate(cases.toy_model, (B, 4, Hh, 64), 1.0, first_endpoint=e0, second_endpoint=e1, return_tensor=True,
                               condition=cond, sampler="ddpm")
    g["interp_init"], g["interp_last"] = init.numpy(), imgs[-1].numpy()
    S = ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=2, height=Hh, max_width=96, noise_strategy="non_repeat")
    S.respace(list(np.linspace(0, 999, 6, dtype=np.int32)))
    big = cases.randn((2, 4, Hh, 96), 95)
    imgs, init = S.sample(cases.toy_model, (2, 4, Hh, 40), return_tensor=True, condition=cond[:2], sampler="ddim", initial_noise=big)
    g["nonrep_init"], g["nonrep_last"] = init.numpy(), imgs[-1].numpy()
    S = ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B, height=Hh)
    x0, nz = cases.randn((B, 4, Hh, 40), 96), cases.randn((B, 4, Hh, 40), 97)
    g["qsample_t"] = np.array([3, 700, 999])
    g["qsample_out"] = S.q_sample(x0, torch.tensor([3, 700, 999]), noise=nz).numpy()
    mask = (cases.randn((B, 4, Hh, 40), 98) > 0).float()
    guide = cases.randn((B, 4, Hh, 64), 99) * 0.5
    S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B, height=Hh), draws)
    S.activate_classifier_free_guidance(3, uncond)
    S.respace(list(np.linspace(0, 999, 6, dtype=np.int32)))
    imgs, _ = S.inpaint_sample(cases.toy_model, (B, 4, Hh, 40), 1.0, guide, mask, return_tensor=True, condition=cond, initial_noise=draws[0])
    g["inpaint_cmask_last"] = imgs[-1].numpy()
    np.savez_compressed(os.path.join(OUT, "sampler2.npz"), **g)


@torch.no_grad()

def text_head():
    """tests/golden/text.npz: the reference's ProjectionHead (model/multimodal_model.py:35-47) on unit-norm features, with the
    synthetic head weights of tests/test_oracle_text.py::_head_sd.  `python -m oracle.make_golden --text`."""
    import importlib.util
    import types
    ref_loader.load()
    m = types.ModuleType("model.timbre_encoder_pretrain")
    m.get_timbre_encoder = lambda *a, **k: None
    sys.modules["model.timbre_encoder_pretrain"] = m
    from model.multimodal_model import ProjectionHead
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_tt", os.path.join(here, "tests", "test_oracle_text.py"))
    tt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tt)
    head = ProjectionHead(embedding_dim=64, projection_dim=64, dropout=0.1, num_layers=2).eval()
    sd = tt._head_sd()
    head.load_state_dict({k.replace("text_projection.", ""): v for k, v in sd.items()})
    y = torch.nn.functional.normalize(torch.randn((5, 64), generator=torch.Generator().manual_seed(2)), dim=-1)
    with torch.no_grad():
        out = head(y)
    np.savez(os.path.join(here, "tests", "golden", "text.npz"), features=y.numpy(), head_out=out.numpy())



VARIANTS = ("small_family_w16", "small_notime_w16", "small_nocond_w16", "small_cat_nocond_w16")


def variants(ref):
    """tests/golden/variants.npz: the reference U-Net's own outputs for condition_type="instrument_family", with_time_emb=False and
    condition=None (diffusion_components.py:155-168; diffusion.py:107-109,199-202,211).  `python -m oracle.make_golden --variants`."""
    out = {}
    for name in VARIANTS:
        cfg, sd, x, t, cond = cases.unet_case(name)
        net = ref.ConditionedUnet(**cfg).eval()
        net.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out[f"{name}_eps"] = net(x, t, cond).numpy()
    np.savez(os.path.join(OUT, "variants.npz"), **out)


def main():
    torch.set_num_threads(8)
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    if "--variants" in sys.argv:
        variants(ref)
        print("variants.npz", os.path.getsize(os.path.join(OUT, "variants.npz")))
        return
    if "--sampler2" in sys.argv:
        sampler2(ref)
        print("sampler2.npz", os.path.getsize(os.path.join(OUT, "sampler2.npz")))
        return
    if "--headline" in sys.argv:
        headline(ref)
        print("headline.npz", os.path.getsize(os.path.join(OUT, "headline.npz")))
        return
    extra(ref)
    if "--text" in sys.argv:
        text_head()
        return
    if "--extra" in sys.argv:
        print("extra.npz", os.path.getsize(os.path.join(OUT, "extra.npz")))
        return

    # ---- sampler ---------------------------------------------------------------------------
    g = {}
    S = ref.DiffSynthSampler(1000, device="cpu", mute=True)
    g["betas_1000"], g["ac_1000"] = S.betas, S.alphas_cumprod
    for steps in (10, 20):
        S = ref.DiffSynthSampler(1000, device="cpu", mute=True)
        S.respace(list(np.linspace(0, 999, steps, dtype=np.int32)))
        g[f"betas_{steps}"], g[f"ac_{steps}"], g[f"acp_{steps}"] = S.betas, S.alphas_cumprod, S.alphas_cumprod_prev
        g[f"map_{steps}"] = np.array(S.timestep_map)
    S = ref.DiffSynthSampler(1000, device="cpu", mute=True, height=2, channels=1, max_batchsize=2)
    base = cases.randn((2, 1, 2, 64), 5)
    for w in (17, 24, 63, 64, 65, 100, 144, 200):
        n, pts = S.get_deterministic_noise_tensor_repeat(2, w, reference_noise=base)
        g[f"layout_{w}"], g[f"layout_pts_{w}"] = n.numpy(), np.array(pts)
    # full loops on a toy model: ddim / ddpm with CFG, CFG=1, img-guided, inpaint (fixed mask)
    B, Wd = 3, 40
    draws = cases.randn((12, B, 4, 128, 64), 6)
    cond, uncond = W.synthetic_conditions(B, 16, seed=77)
    for name, kw in (("ddim", dict(sampler="ddim")), ("ddpm", dict(sampler="ddpm"))):
        S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B), draws)
        S.activate_classifier_free_guidance(6, uncond)
        S.respace(list(np.linspace(0, 999, 8, dtype=np.int32)))
        imgs, init = S.sample(cases.toy_model, (B, 4, 128, Wd), return_tensor=True, condition=cond,
                              initial_noise=draws[0], **kw)
        g[f"loop_{name}_last"], g[f"loop_{name}_mid"] = imgs[-1].numpy(), imgs[3].numpy()
        assert len(imgs) == 9
    S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B), draws)
    S.respace(list(np.linspace(0, 999, 8, dtype=np.int32)))
    imgs, _ = S.sample(cases.toy_model, (B, 4, 128, Wd), return_tensor=True, condition=cond, initial_noise=draws[0])
    g["loop_nocfg_last"] = imgs[-1].numpy()
    guide = cases.randn((B, 4, 128, 64), 8) * 0.5
    S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B), draws)
    S.activate_classifier_free_guidance(6, uncond)
    S.respace(list(np.linspace(0, 999, int(8 / 0.7), dtype=np.int32)))
    # img_guided_sample requires guide width == shape width (:580) and the repeat layout requires
    # guide width == train_width (:113): guided sampling only exists at W = 64.
    imgs, _ = S.img_guided_sample(cases.toy_model, (B, 4, 128, 64), 0.7, guide, return_tensor=True,
                                  condition=cond, initial_noise=draws[0])
    g["loop_guided_first"], g["loop_guided_last"], g["loop_guided_len"] = imgs[0].numpy(), imgs[-1].numpy(), np.array(len(imgs))
    mask = (cases.randn((B, 1, 128, Wd), 9) > 0).float()
    S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B), draws)
    S.activate_classifier_free_guidance(6, uncond)
    S.respace(list(np.linspace(0, 999, 8, dtype=np.int32)))
    imgs, _ = S.inpaint_sample(cases.toy_model, (B, 4, 128, Wd), 1.0, guide, mask, return_tensor=True,
                               condition=cond, initial_noise=draws[0])
    g["loop_inpaint_last"] = imgs[-1].numpy()
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), **g)

    # ---- U-Net -----------------------------------------------------------------------------
    g = {}
    for name in ("deployed_w64", "deployed_w24", "small_w16"):
        cfg, sd, x, t, cond = cases.unet_case(name)
        net = ref.ConditionedUnet(**cfg).eval()
        net.load_state_dict(sd, strict=True)
        g[f"{name}_eps"] = net(x, t, cond).numpy()
        if name == "deployed_w64":      # per-layer (mean, rms) to localise a mismatch
            feats = {}
            hooks = []
            for mod_name, mod in net.named_modules():
                parts = mod_name.split(".")
                if (parts[0] in ("downs", "ups") and len(parts) == 3) or \
                   (parts[0] in ("mid_left", "mid_right", "mid_mid", "final_conv") and len(parts) == 2) or mod_name == "init_conv":
                    hooks.append(mod.register_forward_hook(
                        lambda m, i, o, n=mod_name: feats.__setitem__(n, stats(o))))
            net(x, t, cond)
            for k, v in feats.items():
                g[f"{name}_tap_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "unet.npz"), **g)

    # ---- VQGAN -----------------------------------------------------------------------------
    g = {}
    vq = ref.VQGAN(**W.VQGAN_DEPLOYED).eval()
    sd = W.vqgan_random_state_dict(seed=1)
    vq.load_state_dict(sd, strict=True)
    lat = cases.vq_latents()
    q, _, _ = vq._vq_vae(lat)
    flat = lat.permute(0, 2, 3, 1).reshape(-1, 4)
    cb = sd["_vq_vae._embedding.weight"]
    d = (torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(cb ** 2, dim=1) - 2 * torch.matmul(flat, cb.t()))
    g["vq_idx"] = torch.argmin(d, dim=1).numpy().astype(np.int16)        # VQGAN.py:107-112
    g["vq_q"] = q.numpy()
    dec = vq._decoder(q)
    g["dec_sub"] = dec.flatten()[::7].numpy()
    g["dec_stats"] = stats(dec)
    spec = torch.from_numpy(np.stack([_spec(ref)]))
    enc = vq._encoder(spec)
    g["enc_lat"] = enc.numpy()
    np.savez_compressed(os.path.join(OUT, "vqgan.npz"), **g)

    # ---- STFT+ codec (tools.py) ------------------------------------------------------------
    g = {}
    e = cases.randn((3, 512, 6), 41).numpy()
    e[0] = np.abs(e[0])
    D = ref.tools.decode_stft(e)
    g["decode_in"], g["decode_out"] = e, D
    g["depad_out"] = ref.tools.depad_STFT(D)
    Dc = (cases.randn((513, 5), 42) + 1j * cases.randn((513, 5), 43)).numpy()
    g["encode_in"], g["pad_out"] = Dc, ref.tools.pad_STFT(Dc, 8)
    g["encode_out"] = ref.tools.encode_stft(ref.tools.pad_STFT(Dc, 8))
    np.savez_compressed(os.path.join(OUT, "codec.npz"), **g)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def _spec(ref):
    from oracle import ds_oracle as O
    return O.waveform_to_spectrogram(cases.synthetic_wave())


if __name__ == "__main__":
    main()

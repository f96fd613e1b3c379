"""Seeded inputs shared by oracle/make_golden.py (which runs the reference on them) and the
tests (which run the oracle / CUDA path on them).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import numpy as np
import torch

from diffusynth_b200 import weights as W

SMALL_UNET = dict(in_dim=4, down_dims=[32, 32, 64], up_dims=[64, 64, 32], mid_depth=2,
                  attn_type="linear_add", condition_type="natural_language_prompt", label_emb_dim=64)


SMALL_UNET_CAT = dict(SMALL_UNET, attn_type="linear_cat")
SMALL_UNET_RESNET = dict(SMALL_UNET, use_convnext=False, resnet_block_groups=8)
SMALL_UNET_FAMILY = dict(SMALL_UNET, condition_type="instrument_family", n_label_class=11)      # integer labels through nn.Embedding
SMALL_UNET_NOTIME = dict(SMALL_UNET, with_time_emb=False)


def randn(shape, seed):
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    return torch.randn(shape, generator=g)


def unet_case(name: str):
    """-> (cfg, state_dict, x, t, cond)."""
    if name == "deployed_w64":
        cfg, B, Wd = W.UNET_DEPLOYED, 2, 64
    elif name == "deployed_w24":
        cfg, B, Wd = W.UNET_DEPLOYED, 1, 24
    elif name == "deployed_w28":        # 28 -> 14 -> 7 -> 3: the 7-wide level is odd (pad_to_match pads the upsampled 6 to 7)
        cfg, B, Wd = W.UNET_DEPLOYED, 1, 28
    elif name == "small_w16":
        cfg, B, Wd = SMALL_UNET, 2, 16
    elif name == "small_w10":           # 10 -> 5 -> 2
        cfg, B, Wd = SMALL_UNET, 2, 10
    elif name == "small_resnet_w16":    # use_convnext=False: ResnetBlock (conv3x3 -> GroupNorm(8) -> SiLU) x 2
        cfg, B, Wd = SMALL_UNET_RESNET, 2, 16
    elif name == "small_cat_w16":       # attn_type="linear_cat" (LinearCrossAttention: the condition is an extra key / value token)
        cfg, B, Wd = SMALL_UNET_CAT, 2, 16
    elif name == "small_family_w16":    # condition_type="instrument_family" (diffusion_components.py:160-161)
        cfg, B, Wd = SMALL_UNET_FAMILY, 2, 16
    elif name == "small_notime_w16":    # with_time_emb=False (diffusion.py:107-109)
        cfg, B, Wd = SMALL_UNET_NOTIME, 2, 16
    elif name == "small_nocond_w16":    # condition=None (diffusion.py:199-202)
        cfg, B, Wd = SMALL_UNET, 2, 16
    elif name == "small_cat_nocond_w16":    # condition=None with LinearCrossAttention: no extra token (diffusion_components.py:195-200)
        cfg, B, Wd = SMALL_UNET_CAT, 2, 16
    else:
        raise KeyError(name)
    sd = W.unet_random_state_dict(cfg, seed=0)
    H = 128 if cfg is W.UNET_DEPLOYED else 32
    x = randn((B, 4, H, Wd), 11) * 1.5
    t = torch.tensor([947, 52][:B], dtype=torch.long)
    cond = randn((B, W.unet_config(**cfg)["label_emb_dim"]), 12)
    if name == "small_family_w16":
        cond = torch.tensor([7, 2][:B], dtype=torch.long)
    elif name in ("small_nocond_w16", "small_cat_nocond_w16"):
        cond = None
    return cfg, sd, x, t, cond


def toy_model(x, t, cond):
    """A cheap stand-in eps-predictor for sampler-logic goldens (any callable works)."""
    return 0.3 * torch.tanh(x) + 0.05 * torch.sin(t.float() / 100.0).view(-1, 1, 1, 1) \
        + 0.1 * cond.mean(dim=1).view(-1, 1, 1, 1)


def vq_latents(B=1, seed=21):
    return randn((B, 4, 128, 64), seed) * 0.9


def synthetic_wave(n=65280, seed=31, sr=16000):
    """A decaying harmonic note plus a little noise, peak-normalised (stands in for a preset WAV)."""
    g = np.random.default_rng(seed)
    t = np.arange(n) / sr
    y = sum(np.sin(2 * np.pi * 220.0 * k * t + g.uniform(0, 6.28)) / k for k in range(1, 9))
    y = y * np.exp(-t * 1.2) + 0.01 * g.standard_normal(n)
    return (y / np.abs(y).max()).astype(np.float64)


def spec_representation(B=2, T=12, seed=51):
    """A spectral representation [B, 3, 512, T] like the decoder's output: channel 0 = softplus-like (> 0),
    channels 1/2 = tanh-like (not unit-norm)."""
    e = randn((B, 3, 512, T), seed)
    e[:, 0] = torch.nn.functional.softplus(1.5 * e[:, 0] - 1.0)
    e[:, 1:] = torch.tanh(e[:, 1:])
    return e


def small_latents(B=2, H=16, Wd=8, seed=52):
    return randn((B, 4, H, Wd), seed) * 2.0


HEADLINE_SEEDS = dict(ddim20=101, ddpm20=102, modify20=103)


def headline_inputs(name: str):
    """Seeded inputs of the headline-configuration runs (shared with tests/test_gpu_headline.py and test_oracle_golden.py):
    deployed U-Net, B = 2, CFG 6, 20 steps; "modify20" = BASELINE config 5 (strength 0.7 -> 28 respaced steps, 19 U-Net steps)."""
    B = 2
    cond, uncond = W.synthetic_conditions(B, 512)
    if name == "modify20":
        n_steps = int(20 / 0.7)
        draws = W.host_noise(HEADLINE_SEEDS[name], 1 + int(n_steps * 0.7), B)
        guide = vq_latents(B=1, seed=23).repeat(B, 1, 1, 1)
        return dict(B=B, cond=cond, uncond=uncond, draws=draws, n_steps=n_steps, strength=0.7, guide=guide, sampler="ddim")
    draws = W.host_noise(HEADLINE_SEEDS[name], 21, B)
    return dict(B=B, cond=cond, uncond=uncond, draws=draws, n_steps=20, strength=1.0, guide=None,
                sampler="ddpm" if name == "ddpm20" else "ddim")


def headline_digest(imgs):
    """What the fixture keeps of a list of per-step latents: every 64th value of every step, per-step (mean, rms, absmax),
    and the final latent in full."""
    sub = np.stack([im.flatten()[::64].numpy() for im in imgs])
    st = np.stack([np.array([im.double().mean().item(), im.double().pow(2).mean().sqrt().item(), im.abs().max().item()]) for im in imgs])
    return sub.astype(np.float32), st, imgs[-1].numpy()

"""Live pinning of the oracle (and of the parameter inventory) against the unmodified
reference imported from /root/reference.  Skipped where the reference is absent (GPU box);
the committed golden vectors (tests/test_oracle_golden.py) carry the same evidence there."""
import numpy as np
import pytest
import torch

from diffusynth_b200 import weights as W
from oracle import cases, ds_oracle as O, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def test_param_inventory_matches_reference(ref):
    net = ref.ConditionedUnet(**W.UNET_DEPLOYED)
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == W.unet_param_spec(W.UNET_DEPLOYED)
    vq = ref.VQGAN(**W.VQGAN_DEPLOYED)
    assert [(k, tuple(v.shape)) for k, v in vq.state_dict().items()] == W.vqgan_param_spec(W.VQGAN_DEPLOYED)


@torch.no_grad()
def test_unet_small_live(ref):
    cfg, sd, x, t, cond = cases.unet_case("small_w16")
    net = ref.ConditionedUnet(**cfg).eval()
    net.load_state_dict(sd, strict=True)
    a, b = net(x, t, cond), O.unet_forward(sd, x, t, cond)
    assert (a - b).norm() / a.norm() < 2e-6


@torch.no_grad()
@pytest.mark.parametrize("name", ["small_family_w16", "small_notime_w16", "small_nocond_w16", "small_cat_nocond_w16"])
def test_unet_conditioning_variants_live(ref, name):
    cfg, sd, x, t, cond = cases.unet_case(name)
    net = ref.ConditionedUnet(**cfg).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == W.unet_param_spec(cfg)
    net.load_state_dict(sd, strict=True)
    a, b = net(x, t, cond), O.unet_forward(sd, x, t, cond)
    assert (a - b).norm() / a.norm() < 2e-6


@torch.no_grad()
def test_full_sample_live(ref):
    """Reference sample() with CFG on the small U-Net vs the oracle loop, same host noise."""
    cfg, sd, _, _, _ = cases.unet_case("small_w16")
    net = ref.ConditionedUnet(**cfg).eval()
    net.load_state_dict(sd, strict=True)
    B = 2
    draws = cases.randn((6, B, 4, 32, 64), 3)
    cond, uncond = W.synthetic_conditions(B, 64, seed=5)
    S = ref_loader.feed_noise(ref.DiffSynthSampler(1000, device="cpu", mute=True, max_batchsize=B, height=32), draws)
    S.activate_classifier_free_guidance(6, uncond)
    S.respace(list(np.linspace(0, 999, 5, dtype=np.int32)))
    imgs, _ = S.sample(net, (B, 4, 32, 32), return_tensor=True, condition=cond, initial_noise=draws[0], sampler="ddpm")
    s = O.Schedule(1000)
    s.respace(list(np.linspace(0, 999, 5, dtype=np.int32)))
    mine = O.sample_loop(lambda x, t, c: O.unet_forward(sd, x, t, c), s, (B, 4, 32, 32), cond, uncond, 6, draws,
                         sampler="ddpm")
    assert len(imgs) == len(mine) == 6
    for a, b in zip(imgs, mine):
        assert (a - b).norm() / a.norm() < 1e-5


@torch.no_grad()
def test_vq_live(ref):
    vq = ref.VQGAN(**W.VQGAN_DEPLOYED).eval()
    sd = W.vqgan_random_state_dict(seed=1)
    vq.load_state_dict(sd, strict=True)
    lat = cases.randn((1, 4, 32, 64), 99)
    q_ref, _, _ = vq._vq_vae(lat)
    q, idx = O.vq_quantize(lat, sd["_vq_vae._embedding.weight"])
    flat = lat.permute(0, 2, 3, 1).reshape(-1, 4)
    cb = sd["_vq_vae._embedding.weight"]
    d = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(cb ** 2, dim=1) - 2 * torch.matmul(flat, cb.t())
    assert np.array_equal(O.vq_distances(flat.numpy(), cb.numpy()), d.numpy()), "distance bits differ from torch CPU"
    assert torch.equal(idx, torch.argmin(d, dim=1)) and torch.equal(q, q_ref)

"""Sampler surfaces beyond sample(): interpolate(), the "non_repeat" noise strategy, q_sample with per-sample timesteps, inpainting
with a per-channel [B,C,H,W] mask (the shape inpaint_with_text.py:229-231 passes) on both the generic and the CUDA-graph loop,
a [1,1,H,W] mask broadcast over the batch, and the graph cache after a weight reload -- against goldens minted from the unmodified
reference (tests/golden/sampler2.npz, oracle/make_golden.py --sampler2) and the oracle.  fp32 kernels: 1e-5."""
import numpy as np
import pytest
import torch

from diffusynth_b200 import weights as W
from oracle import cases, ds_oracle as O
from tests.gpu_util import rel

pytestmark = pytest.mark.gpu
B, Hh = 3, 16


def _sampler(draws=None, cfg=True, **kw):
    from diffusynth_b200 import DiffSynthSampler
    cond, uncond = W.synthetic_conditions(B, 16, seed=91)
    s = DiffSynthSampler(1000, device="cuda", mute=True, height=Hh, **kw)
    if cfg:
        s.activate_classifier_free_guidance(3, uncond.cuda())
    s.respace(list(np.linspace(0, 999, 6, dtype=np.int32)))
    if draws is not None:
        s.noise_feed = draws[1:]
    return s, cond


def test_interpolate_between_endpoints(golden):
    g = golden["sampler2"]
    draws = cases.randn((12, B, 4, Hh, 64), 92)
    e0, e1 = cases.randn((4, Hh, 64), 93), cases.randn((4, Hh, 64), 94)
    s, cond = _sampler(draws, max_batchsize=B)
    imgs, init = s.interpolate(cases.toy_model, (B, 4, Hh, 64), 1.0, first_endpoint=e0.cuda(), second_endpoint=e1.cuda(),
                               return_tensor=True, condition=cond.cuda(), sampler="ddpm")
    assert rel(init, torch.from_numpy(g["interp_init"])) < 1e-6 and rel(imgs[-1], torch.from_numpy(g["interp_last"])) < 1e-5


def test_non_repeat_noise_strategy(golden):
    g = golden["sampler2"]
    s, cond = _sampler(None, cfg=False, max_batchsize=2, max_width=96, noise_strategy="non_repeat")
    big = cases.randn((2, 4, Hh, 96), 95)
    imgs, init = s.sample(cases.toy_model, (2, 4, Hh, 40), return_tensor=True, condition=cond[:2].cuda(), sampler="ddim", initial_noise=big.cuda())
    assert torch.equal(init.cpu(), torch.from_numpy(g["nonrep_init"])) and rel(imgs[-1], torch.from_numpy(g["nonrep_last"])) < 1e-5
    with pytest.raises(AssertionError):          # the reference asserts the full max_width shape (:71)
        s.sample(cases.toy_model, (2, 4, Hh, 40), condition=cond[:2].cuda(), initial_noise=big[..., :64].cuda())


def test_q_sample_per_sample_timesteps(golden):
    g = golden["sampler2"]
    s, _ = _sampler(None, cfg=False, max_batchsize=B)
    s2 = __import__("diffusynth_b200").DiffSynthSampler(1000, device="cuda", mute=True, height=Hh, max_batchsize=B)     # un-respaced, like the fixture
    x0, nz = cases.randn((B, 4, Hh, 40), 96), cases.randn((B, 4, Hh, 40), 97)
    out = s2.q_sample(x0.cuda(), torch.tensor(g["qsample_t"]).cuda(), noise=nz.cuda())
    assert rel(out, torch.from_numpy(g["qsample_out"])) < 1e-6
    one = s2.q_sample(x0.cuda(), torch.full((B,), 700).cuda(), noise=nz.cuda())
    assert torch.equal(one[1], out[1])


@pytest.mark.parametrize("path", ["generic", "graph"])
def test_inpaint_with_per_channel_mask(path, golden):
    """mask [B,C,H,W] with different values per channel: generic loop against the reference's own result on the toy model; graph
    loop (B200 U-Net) against the oracle loop, together with a [1,1,H,W] mask shared by the whole batch."""
    g = golden["sampler2"]
    draws = cases.randn((12, B, 4, Hh, 64), 92)
    mask = (cases.randn((B, 4, Hh, 40), 98) > 0).float()
    guide = cases.randn((B, 4, Hh, 64), 99) * 0.5
    if path == "generic":
        s, cond = _sampler(draws, max_batchsize=B)
        imgs, _ = s.inpaint_sample(cases.toy_model, (B, 4, Hh, 40), 1.0, guide.cuda(), mask.cuda(), return_tensor=True, condition=cond.cuda(),
                                   initial_noise=draws[0].cuda())
        assert rel(imgs[-1], torch.from_numpy(g["inpaint_cmask_last"])) < 1e-5
        return
    from diffusynth_b200 import ConditionedUnet, DiffSynthSampler
    cfg, sd, _, _, _ = cases.unet_case("small_w16")
    net = ConditionedUnet(**{k: v for k, v in cfg.items() if k not in ("out_dim", "time_dim")}, device="cuda")
    net.load_state_dict(sd)
    H2 = 32
    draws = cases.randn((8, B, 4, H2, 64), 192)
    guide = cases.randn((B, 4, H2, 64), 199) * 0.5
    cond, uncond = W.synthetic_conditions(B, 64, seed=300)
    model = lambda x, t, c: O.unet_forward(sd, x, t, c)
    for mk in ((cases.randn((B, 4, H2, 40), 198) > 0).float(), (cases.randn((1, 1, H2, 40), 197) > 0).float()):
        s = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=B, height=H2)
        s.activate_classifier_free_guidance(4, uncond.cuda())
        s.respace(list(np.linspace(0, 999, 4, dtype=np.int32)))
        s.noise_feed = draws[1:]
        imgs, _ = s.inpaint_sample(net, (B, 4, H2, 40), 1.0, guide.cuda(), mk.cuda(), return_tensor=True, condition=cond.cuda(),
                                   initial_noise=draws[0].cuda())
        assert s.last_graph_launches > 0
        sch = O.Schedule(1000)
        sch.respace(list(np.linspace(0, 999, 4, dtype=np.int32)))
        with torch.no_grad():
            ref = O.sample_loop(model, sch, (B, 4, H2, 40), cond, uncond, 4, draws, guide=guide, mask=mk, inpaint=True)
        errs = [rel(a, b) for a, b in zip(imgs, ref)]
        print(f"\nmask {tuple(mk.shape)}: latents rel-L2 per step {['%.1e' % e for e in errs]}")
        assert max(errs) < 1e-2


def test_graph_cache_follows_weight_reload():
    """Reloading a checkpoint into the same model object must not keep sampling with the graph captured on the old weights."""
    from diffusynth_b200 import ConditionedUnet, DiffSynthSampler
    cfg, sd, _, _, _ = cases.unet_case("small_w16")
    net = ConditionedUnet(**{k: v for k, v in cfg.items() if k not in ("out_dim", "time_dim")}, device="cuda")
    net.load_state_dict(sd)
    H2 = 32
    draws = cases.randn((6, 2, 4, H2, 64), 292)
    cond, uncond = W.synthetic_conditions(2, 64, seed=300)
    s = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=2, height=H2)
    s.activate_classifier_free_guidance(4, uncond.cuda())
    s.respace(list(np.linspace(0, 999, 3, dtype=np.int32)))

    def run():
        s.noise_feed = draws[1:]
        return s.sample(net, (2, 4, H2, 32), return_tensor=True, condition=cond.cuda(), initial_noise=draws[0].cuda())[0][-1]

    a = run()
    sd2 = W.unet_random_state_dict(cfg, seed=5)
    net.load_state_dict(sd2)
    b = run()
    assert len(s._graphs) == 1 and rel(a, b) > 1e-3            # the old loop was retired, the result follows the new weights
    fresh = ConditionedUnet(**{k: v for k, v in cfg.items() if k not in ("out_dim", "time_dim")}, device="cuda")
    fresh.load_state_dict(sd2)
    s2 = DiffSynthSampler(1000, device="cuda", mute=True, max_batchsize=2, height=H2)
    s2.activate_classifier_free_guidance(4, uncond.cuda())
    s2.respace(list(np.linspace(0, 999, 3, dtype=np.int32)))
    s2.noise_feed = draws[1:]
    c = s2.sample(fresh, (2, 4, H2, 32), return_tensor=True, condition=cond.cuda(), initial_noise=draws[0].cuda())[0][-1]
    assert torch.equal(b, c)

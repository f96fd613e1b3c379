"""Module-level C ABI (ds_unet_*, ds_vqgan_*, ds_sample_graph_*): the library's own plans, packing and graph against the
operator-level Python plans (which the per-layer parity tests pin to the oracle), and a U-Net forward driven by a plain C program
that links nothing but the shared library."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _unet(perturb=True):
    from diffusynth_b200 import ConditionedUnet, weights as W
    net = ConditionedUnet(**{k: v for k, v in W.UNET_DEPLOYED.items() if k not in ("out_dim", "time_dim")}, device="cuda")
    net.load_state_dict(W.unet_random_state_dict(seed=3, perturb_norm=perturb))
    return net


@pytest.mark.parametrize("shape", [(2, 128, 64), (3, 128, 28), (2, 32, 24)])
def test_unet_forward_engine_equals_operator_plan(shape):
    """ds_unet_forward (packing + plan inside the library) reproduces the operator-level plan bit for bit: same kernels, same
    arguments, same tables (odd levels included: W = 28 -> 14 -> 7)."""
    N, H, Wd = shape
    net = _unet()
    assert net._engine is not None
    g = torch.Generator().manual_seed(5)
    x = torch.randn((N, 4, H, Wd), generator=g).cuda()
    t = torch.tensor([17, 500, 999][:N]).cuda()
    c = torch.randn((N, 512), generator=g).cuda()
    e_eng = net(x, t, c)
    net.use_engine = False
    e_py = net(x, t, c)
    torch.cuda.synchronize()
    assert torch.isfinite(e_eng).all()
    r = rel(e_eng, e_py)
    print(f"\nengine vs operator plan {shape}: rel-L2 {r:.2e}, identical {bool(torch.equal(e_eng, e_py))}")
    assert torch.equal(e_eng, e_py)


def test_vqgan_engine_equals_operator_plan():
    from diffusynth_b200 import VQGAN, weights as W
    vq = VQGAN(**W.VQGAN_DEPLOYED, device="cuda")
    vq.load_state_dict(W.vqgan_random_state_dict(seed=4))
    g = torch.Generator().manual_seed(6)
    lat = torch.randn((2, 4, 128, 64), generator=g).cuda()
    q, _, _ = vq._vq_vae(lat)
    spec_e = vq._decoder(q)
    z_e = vq._encoder(spec_e)
    vq._decoder.use_engine = vq._encoder.use_engine = False
    spec_p = vq._decoder(q)
    z_p = vq._encoder(spec_e)
    r1, r2 = rel(spec_e, spec_p), rel(z_e, z_p)
    print(f"\ndecoder engine vs plan {r1:.2e}, encoder {r2:.2e}")
    assert tuple(spec_e.shape) == (2, 3, 512, 256) and tuple(z_e.shape) == (2, 4, 128, 64)
    assert torch.equal(spec_e, spec_p) and torch.equal(z_e, z_p)
    # quantiser through the handle's codebook
    from diffusynth_b200 import _lib
    out = torch.empty_like(lat)
    idx = torch.empty((2 * 128 * 64,), dtype=torch.long, device="cuda")
    _lib.check(_lib.load().ds_vqgan_quantize(vq._engine.h, lat.data_ptr(), out.data_ptr(), idx.data_ptr(), 2, 128 * 64,
                                            torch.cuda.current_stream().cuda_stream))
    assert torch.equal(out, q) and torch.equal(idx, vq._vq_vae.last_indices)


@pytest.mark.parametrize("sampler", ["ddim", "ddpm"])
def test_sample_graph_engine_equals_python_graph(sampler):
    """ds_sample_graph_build / _run (capture inside the library, tail included) against the torch.cuda.graph loop over the
    operator-level plan: same latents, quantised latents, spectrograms and waveforms."""
    from diffusynth_b200 import TextToTimbre, weights as W
    B, steps = 2, 4
    cond, uncond = W.synthetic_conditions(B, 512)
    feed = W.host_noise(11, steps + 1, B)
    outs = []
    for use in (True, False):
        pipe = TextToTimbre.random_init(device="cuda", seed=0)
        pipe.unet.use_engine = use
        o = pipe.generate(cond.cuda(), uncond.cuda(), steps=steps, cfg_scale=6.0, width=64, sampler=sampler, noise_feed=feed)
        s = pipe._samplers[(B, steps)]
        loop = next(iter(s._graphs.values()))
        assert (loop.sgraph is not None) == use
        outs.append(o)
    a, b = outs
    rs = [rel(a.latents, b.latents), rel(a.spectrograms, b.spectrograms), rel(a.waveforms, b.waveforms)]
    print(f"\n{sampler}: library graph vs python graph: latents {rs[0]:.2e}, spectrograms {rs[1]:.2e}, waveforms {rs[2]:.2e}")
    assert torch.isfinite(a.waveforms).all()
    assert torch.equal(a.latents, b.latents) and torch.equal(a.quantized, b.quantized) and torch.equal(a.waveforms, b.waveforms)


def test_unet_forward_from_plain_c(tmp_path):
    """A C host (tests/abi_unet_forward.c: no Python, no torch) loads the parameters by their reference names and calls
    ds_unet_forward; its epsilon equals the Python host's."""
    from diffusynth_b200 import weights as W
    cfg = W.unet_config(down_dims=[32, 32, 64], up_dims=[64, 64, 32], label_emb_dim=64)
    from diffusynth_b200 import ConditionedUnet
    net = ConditionedUnet(in_dim=4, down_dims=cfg["down_dims"], up_dims=cfg["up_dims"], attn_type="linear_add",
                          condition_type="natural_language_prompt", label_emb_dim=64, device="cuda")
    sd = W.random_state_dict(W.unet_param_spec(cfg), seed=8)
    net.load_state_dict(sd)
    N, H, Wd = 2, 32, 16
    g = torch.Generator().manual_seed(9)
    x, t, c = torch.randn((N, 4, H, Wd), generator=g), torch.tensor([3, 700]), torch.randn((N, 64), generator=g)
    ref = net(x.cuda(), t.cuda(), c.cuda()).cpu()
    wfile, ifile, ofile = tmp_path / "weights.bin", tmp_path / "inputs.bin", tmp_path / "eps.bin"
    with open(wfile, "wb") as f:
        for name, ten in sd.items():
            nb = name.encode()
            f.write(struct.pack("<i", len(nb)) + nb + struct.pack("<i", ten.dim()) + struct.pack(f"<{ten.dim()}q", *ten.shape))
            f.write(ten.float().contiguous().numpy().tobytes())
        f.write(struct.pack("<i", 0))
    with open(ifile, "wb") as f:
        f.write(struct.pack("<3i", N, H, Wd) + x.numpy().tobytes() + t.numpy().astype(np.int64).tobytes() + c.numpy().tobytes())
    exe = tmp_path / "abi_unet_forward"
    lib_dir = os.path.join(ROOT, "diffusynth_b200")
    subprocess.check_call(["gcc", "-O1", "-o", str(exe), os.path.join(ROOT, "tests", "abi_unet_forward.c"), "-I", os.path.join(ROOT, "include"),
                           "-I", "/usr/local/cuda/include", "-L", lib_dir, "-ldiffusynth_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                           f"-Wl,-rpath,{lib_dir}", "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([str(exe), str(wfile), str(ifile), str(ofile)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    eps = torch.from_numpy(np.fromfile(ofile, dtype=np.float32).reshape(N, 4, H, Wd))
    r = rel(eps, ref)
    print(f"\nC host vs Python host: rel-L2 {r:.2e}, identical {bool(torch.equal(eps, ref))}")
    assert torch.equal(eps, ref)

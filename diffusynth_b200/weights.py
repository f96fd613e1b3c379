"""Parameter inventory (names + shapes) of the two networks on the sampling path,
and a deterministic random initialiser for them.

The names are exactly the reference's ``state_dict`` keys, so a trained
checkpoint (``torch.load(...)['model_state_dict']``, reference
``model/diffusion.py:371-374`` and ``model/VQGAN.py:581-584``) loads into the
B200 classes unchanged, and so that the very same tensors can be handed to the
reference, the oracle and the CUDA path in the parity tests.

Nothing here reads the reference tree; ``tests/test_oracle_vs_reference.py``
checks this inventory against the reference constructors when the reference is
available.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Tuple

import torch

# Deployed configurations (reference app.py:32-35 and app.py:40).
UNET_DEPLOYED = dict(
    in_dim=4, out_dim=None, down_dims=[96, 96, 192, 384], up_dims=[384, 384, 192, 96],
    mid_depth=3, time_dim=None, convnext_mult=2, attn_type="linear_add",
    condition_type="natural_language_prompt", label_emb_dim=512,
)
VQGAN_DEPLOYED = dict(
    in_channels=3, hidden_channels=[80, 160], embedding_dim=4, out_channels=3, block_depth=2,
    attn_pos=[80, 160], attn_with_skip=True, num_embeddings=8192, commitment_cost=0.25,
    decay=0.99, norm_type="groupnorm", act_type="swish", num_groups=16,
)

ATTN_HEADS = 4        # LinearCrossAttentionAdd default heads  (diffusion_components.py:253)
ATTN_DIM_HEAD = 32    # ... and dim_head
VQ_ATTN_DIM = 32      # VQGAN LinearAttention(current_channel, 1, 32, ...) (VQGAN.py:294,346)

Spec = List[Tuple[str, Tuple[int, ...]]]


def unet_config(**overrides) -> dict:
    cfg = dict(UNET_DEPLOYED)
    cfg.update(overrides)
    if cfg.get("out_dim") is None:
        cfg["out_dim"] = cfg["in_dim"]
    if cfg.get("time_dim") is None:
        cfg["time_dim"] = int(cfg["down_dims"][0] * 4)
    cfg.setdefault("use_convnext", True)
    cfg.setdefault("resnet_block_groups", 8)
    cfg.setdefault("with_time_emb", True)
    cfg.setdefault("n_label_class", 11)
    return cfg


def _convnext(spec: Spec, p: str, dim: int, dim_out: int, mult: int, time_dim) -> None:
    if time_dim is not None:
        spec += [(p + "mlp.1.weight", (dim, time_dim)), (p + "mlp.1.bias", (dim,))]
    spec += [
        (p + "ds_conv.weight", (dim, 1, 7, 7)), (p + "ds_conv.bias", (dim,)),
        (p + "net.0.weight", (dim,)), (p + "net.0.bias", (dim,)),
        (p + "net.1.weight", (dim_out * mult, dim, 3, 3)), (p + "net.1.bias", (dim_out * mult,)),
        (p + "net.3.weight", (dim_out * mult,)), (p + "net.3.bias", (dim_out * mult,)),
        (p + "net.4.weight", (dim_out, dim_out * mult, 3, 3)), (p + "net.4.bias", (dim_out,)),
    ]
    if dim != dim_out:
        spec += [(p + "res_conv.weight", (dim_out, dim, 1, 1)), (p + "res_conv.bias", (dim_out,))]


def _resnet(spec: Spec, p: str, dim: int, dim_out: int, time_dim) -> None:
    """ResnetBlock (use_convnext=False), registration order of diffusion_components.py:82-93."""
    if time_dim is not None:
        spec += [(p + "mlp.1.weight", (dim_out, time_dim)), (p + "mlp.1.bias", (dim_out,))]
    for b, cin in (("block1.", dim), ("block2.", dim_out)):
        spec += [(p + b + "proj.weight", (dim_out, cin, 3, 3)), (p + b + "proj.bias", (dim_out,)),
                 (p + b + "norm.weight", (dim_out,)), (p + b + "norm.bias", (dim_out,))]
    if dim != dim_out:
        spec += [(p + "res_conv.weight", (dim_out, dim, 1, 1)), (p + "res_conv.bias", (dim_out,))]


def _attn(spec: Spec, p: str, dim: int, label_dim: int, attn_type: str = "linear_add") -> None:
    """LinearCrossAttentionAdd (label_key, label_query; diffusion_components.py:252-269) or LinearCrossAttention
    ("linear_cat": label_key, label_value; :171-185), wrapped in Residual(PreNorm(.))."""
    hid = ATTN_HEADS * ATTN_DIM_HEAD
    second = "label_query" if attn_type == "linear_add" else "label_value"
    spec += [
        (p + "fn.fn.to_qkv.weight", (hid * 3, dim, 1, 1)),
        (p + "fn.fn.to_out.0.weight", (dim, hid, 1, 1)), (p + "fn.fn.to_out.0.bias", (dim,)),
        (p + "fn.fn.to_out.1.weight", (dim,)), (p + "fn.fn.to_out.1.bias", (dim,)),
        (p + "fn.fn.label_key.weight", (hid, label_dim)), (p + "fn.fn.label_key.bias", (hid,)),
        (p + f"fn.fn.{second}.weight", (hid, label_dim)), (p + f"fn.fn.{second}.bias", (hid,)),
        (p + "fn.norm.weight", (dim,)), (p + "fn.norm.bias", (dim,)),
    ]


def unet_param_spec(cfg: dict) -> Spec:
    """(name, shape) in the reference's registration order (diffusion.py:62-175)."""
    cfg = unet_config(**cfg)
    dd, ud = cfg["down_dims"], cfg["up_dims"]
    L, td, mult = cfg["label_emb_dim"], cfg["time_dim"], cfg["convnext_mult"]
    assert cfg["attn_type"] in ("linear_add", "linear_cat") and cfg["condition_type"] in ("natural_language_prompt", "instrument_family")
    at = cfg["attn_type"]
    if not cfg["with_time_emb"]:
        td = None                                          # diffusion.py:107-109: no time_mlp, blocks built without their mlp

    def _blk(spec, p, dim, dim_out, mult, time_dim):      # block_klass (diffusion.py:83-87)
        if cfg["use_convnext"]:
            _convnext(spec, p, dim, dim_out, mult, time_dim)
        else:
            _resnet(spec, p, dim, dim_out, time_dim)

    if cfg["condition_type"] == "natural_language_prompt":   # ConditionalEmbedding, diffusion_components.py:155-168
        spec: Spec = [("label_embedding.embedding.weight", (L, L)), ("label_embedding.embedding.bias", (L,))]
    else:                                                     # nn.Embedding(n_label_class + 1, label_emb_dim)
        spec = [("label_embedding.embedding.weight", (cfg["n_label_class"] + 1, L))]      # one extra token (diffusion.py:63-64)
    spec += [("init_conv.weight", (dd[0], cfg["in_dim"], 7, 7)), ("init_conv.bias", (dd[0],))]
    if td is not None:
        spec += [("time_mlp.1.weight", (td, dd[0])), ("time_mlp.1.bias", (td,)),
                 ("time_mlp.3.weight", (td, td)), ("time_mlp.3.bias", (td,))]
    skips = []
    for i, (cin, cout) in enumerate(zip(dd[:-1], dd[1:])):
        p = f"downs.{i}."
        _blk(spec, p + "0.", cin, cout, mult, td)
        _attn(spec, p + "1.", cout, L, at)
        _blk(spec, p + "2.", cout, cout, mult, td)
        _attn(spec, p + "3.", cout, L, at)
        spec += [(p + "4.weight", (cout, cout, 4, 4)), (p + "4.bias", (cout,))]
        skips.append(cout)
    mid = dd[-1]
    # registration order in the reference: downs, ups (empty list first), mid_left, mid_right, mid_mid
    ups_spec: Spec = []
    sk = list(skips)
    for i, (cin, cout) in enumerate(zip(ud[:-1], ud[1:])):
        s = sk.pop()
        p = f"ups.{i}."
        _blk(ups_spec, p + "0.", cin + s, cin, mult, td)
        _attn(ups_spec, p + "1.", cin, L, at)
        ups_spec += [(p + "2.weight", (cin, cin, 4, 4)), (p + "2.bias", (cin,))]
        _blk(ups_spec, p + "3.", cin + s, cout, mult, td)
        _attn(ups_spec, p + "4.", cout, L, at)
        _blk(ups_spec, p + "5.", cout + s, cout, mult, td)
        _attn(ups_spec, p + "6.", cout, L, at)
    spec += ups_spec
    for j in range(cfg["mid_depth"] - 1):
        _blk(spec, f"mid_left.{j}.", mid, mid, mult, td)
    for j in range(cfg["mid_depth"] - 1):
        _blk(spec, f"mid_right.{j}.", mid * 2, mid, mult, td)
    _blk(spec, "mid_mid.0.", mid, mid, mult, td)
    _attn(spec, "mid_mid.1.", mid, L, at)
    _blk(spec, "mid_mid.2.", mid, mid, mult, td)
    _blk(spec, "final_conv.0.", dd[0] + ud[-1], ud[-1], mult, None)
    spec += [("final_conv.1.weight", (cfg["out_dim"], ud[-1], 3, 3)), ("final_conv.1.bias", (cfg["out_dim"],))]
    return spec


def _vq_res(spec: Spec, p: str, cin: int, cout: int) -> None:
    spec += [(p + "norm1.weight", (cin,)), (p + "norm1.bias", (cin,)),
             (p + "conv1.weight", (cout, cin, 3, 3)), (p + "conv1.bias", (cout,)),
             (p + "temb_proj.weight", (cout, 512)), (p + "temb_proj.bias", (cout,))]
    if cin != cout:
        spec += [(p + "nin_shortcut.weight", (cout, cin, 1, 1)), (p + "nin_shortcut.bias", (cout,))]


def _vq_attn(spec: Spec, p: str, c: int, with_skip: bool) -> None:
    spec += [(p + "to_qkv.weight", (3 * VQ_ATTN_DIM, c, 1, 1)),
             (p + "to_out.weight", (c, VQ_ATTN_DIM, 1, 1)), (p + "to_out.bias", (c,))]
    if with_skip:
        spec += [(p + "nin_shortcut.weight", (c, c, 1, 1)), (p + "nin_shortcut.bias", (c,))]


def vqgan_layer_plan(cfg: dict) -> Tuple[list, list]:
    """Layer lists of Encoder (VQGAN.py:278-321) and Decoder (VQGAN.py:332-387):
    entries are (index, kind, c_in, c_out)."""
    hc, depth, attn_pos = list(cfg["hidden_channels"]), cfg["block_depth"], cfg["attn_pos"] or []
    enc, idx = [(0, "down", cfg["in_channels"], hc[0])], 1
    cur = hc[0]

    def enc_blocks():
        nonlocal idx
        for _ in range(depth - 1):
            enc.append((idx, "res", cur, cur)); idx += 1
            if cur in attn_pos:
                enc.append((idx, "attn", cur, cur)); idx += 1

    for i in range(1, len(hc)):
        enc_blocks()
        enc.append((idx, "norm", cur, cur)); idx += 1
        enc.append((idx, "relu", cur, cur)); idx += 1
        enc.append((idx, "down", cur, hc[i])); idx += 1
        cur = hc[i]
    enc_blocks()
    enc.append((idx, "norm", cur, cur)); idx += 1
    enc.append((idx, "relu", cur, cur)); idx += 1
    enc.append((idx, "conv1x1", cur, cfg["embedding_dim"]))

    rc = list(reversed(hc))
    dec, idx = [(0, "conv1x1_nobias", cfg["embedding_dim"], rc[0])], 1
    cur = rc[0]

    def dec_blocks():
        nonlocal idx
        for _ in range(depth - 1):
            if cur in attn_pos:
                dec.append((idx, "attn", cur, cur)); idx += 1
            dec.append((idx, "res", cur, cur)); idx += 1

    dec_blocks()
    for i in range(1, len(rc)):
        dec.append((idx, "norm", cur, cur)); idx += 1
        dec.append((idx, "relu", cur, cur)); idx += 1
        dec.append((idx, "up", cur, rc[i])); idx += 1
        cur = rc[i]
        dec_blocks()
    dec.append((idx, "norm", cur, cur)); idx += 1
    dec.append((idx, "relu", cur, cur)); idx += 1
    dec.append((idx, "up", cur, cur)); idx += 1
    dec.append((idx, "res", cur, cfg["out_channels"]))
    return enc, dec


def vqgan_param_spec(cfg: dict) -> Spec:
    enc, dec = vqgan_layer_plan(cfg)
    spec: Spec = []
    for pre, plan in (("_encoder._layers.", enc), ("_vq_vae", None), ("_decoder._layers.", dec)):
        if plan is None:
            K, D = cfg["num_embeddings"], cfg["embedding_dim"]
            spec += [("_vq_vae._ema_w", (K, D)), ("_vq_vae._ema_cluster_size", (K,)),
                     ("_vq_vae._embedding.weight", (K, D))]
            continue
        for idx, kind, cin, cout in plan:
            p = f"{pre}{idx}."
            if kind == "down":
                spec += [(p + "_conv2d.weight", (cout, cin, 4, 4)), (p + "_conv2d.bias", (cout,))]
            elif kind == "up":   # ConvTranspose2d weight is [in, out, kh, kw]
                spec += [(p + "_conv2d.weight", (cin, cout, 4, 4)), (p + "_conv2d.bias", (cout,))]
            elif kind == "res":
                _vq_res(spec, p, cin, cout)
            elif kind == "attn":
                _vq_attn(spec, p, cin, cfg["attn_with_skip"])
            elif kind == "norm":
                spec += [(p + "weight", (cin,)), (p + "bias", (cin,))]
            elif kind == "conv1x1":
                spec += [(p + "weight", (cout, cin, 1, 1)), (p + "bias", (cout,))]
            elif kind == "conv1x1_nobias":
                spec += [(p + "weight", (cout, cin, 1, 1))]
    return spec


def _is_norm_key(name: str, shape: Tuple[int, ...]) -> bool:
    if len(shape) != 1:
        return False
    stem = name.rsplit(".", 1)[0]
    return stem.endswith(("net.0", "net.3", "to_out.1", "fn.norm", "norm1")) or (
        stem.split(".")[-1].isdigit() and "_layers" in stem)


def random_state_dict(spec: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0,
                      perturb_norm: bool = True, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights (CPU generator, fp32).

    Conv/linear tensors: U(-b, b) with b = gain/sqrt(fan_in) (the torch default
    scale the reference's constructors use).  GroupNorm affine: 1 and 0 as in
    the reference, or, with ``perturb_norm`` (default), 1+0.1*N(0,1) and
    0.1*N(0,1) so that parity tests exercise the affine-folding paths.
    The codebook follows ``normal_()`` (VQGAN.py:88)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    last_fan_in = 1
    for name, shape in spec:
        if name.startswith("_vq_vae."):
            if name.endswith("_ema_cluster_size"):
                t = torch.zeros(shape)
            else:
                t = torch.randn(shape, generator=g)
        elif _is_norm_key(name, shape):
            if name.endswith("weight"):
                t = torch.ones(shape)
                if perturb_norm:
                    t = t + 0.1 * torch.randn(shape, generator=g)
            else:
                t = torch.zeros(shape)
                if perturb_norm:
                    t = 0.1 * torch.randn(shape, generator=g)
        else:
            if len(shape) == 1:      # bias of conv/linear: fan-in of the matching weight (previous entry)
                fan_in = max(1, last_fan_in)
            else:
                fan_in = int(math.prod(shape[1:]))
                is_conv_t = (name.startswith("_decoder") and name.endswith("_conv2d.weight")) or (
                    name.startswith("ups.") and len(name.split(".")) == 3 and len(shape) == 4)
                if is_conv_t:
                    fan_in = shape[0] * 4    # transposed conv: each output sees 2x2 taps of every input channel
                last_fan_in = fan_in
            b = gain / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        sd[name] = t.contiguous()
    return sd



def unet_random_state_dict(cfg: dict | None = None, seed: int = 0, perturb_norm: bool = True):
    return random_state_dict(unet_param_spec(cfg or UNET_DEPLOYED), seed=seed, perturb_norm=perturb_norm)


def vqgan_random_state_dict(cfg: dict | None = None, seed: int = 1, perturb_norm: bool = True):
    return random_state_dict(vqgan_param_spec(cfg or VQGAN_DEPLOYED), seed=seed, perturb_norm=perturb_norm)


def synthetic_conditions(batch: int, dim: int = 512, seed: int = 1000) -> Tuple[torch.Tensor, torch.Tensor]:
    """Synthetic prompt embeddings standing in for CLAP text features:
    cond_i = randn(512; seed+i), uncond = randn(512; seed-1)."""
    conds = []
    for i in range(batch):
        g = torch.Generator(device="cpu"); g.manual_seed(seed + i)
        conds.append(torch.randn(dim, generator=g))
    g = torch.Generator(device="cpu"); g.manual_seed(seed - 1)
    return torch.stack(conds), torch.randn(dim, generator=g)


def host_noise(seed: int, count: int, batch: int, channels: int = 4, height: int = 128,
               train_width: int = 64) -> torch.Tensor:
    """Host-generated noise fed to both implementations: [count, B, C, H, train_width];
    index 0 is the initial noise, 1.. are the per-step draws in loop order
    (reference DiffSynthSampler.py:111 and :340 draw this shape from the global RNG)."""
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    return torch.randn((count, batch, channels, height, train_width), generator=g)

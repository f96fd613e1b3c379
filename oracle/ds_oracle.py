"""CPU ORACLE for the DiffuSynth text-to-timbre sampling path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional torch-fp32 / numpy-f64 on the CPU, what the
reference (WxuanYuan/diffusynth) computes on the sampling path.  It exists so that the
CUDA path in ``diffusynth_b200`` can be checked on a machine where ``/root/reference`` is
not present (the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package never does.

Pinning status (see DESIGN.md "Oracle"):
  * sampler tables / respacing / noise layout / DDIM-DDPM step / loop, U-Net forward,
    VectorQuantizerEMA (eval), VQGAN Decoder and Encoder, decode_stft / encode_stft /
    pad_STFT / depad_STFT: PINNED -- checked against the unmodified reference imported
    from /root/reference (tests/test_oracle_vs_reference.py, runs where the reference
    exists) and against golden vectors minted from the reference by
    oracle/make_golden.py (tests/golden/*.npz, checked everywhere).
  * istft / stft: the arithmetic lives in the third-party ``librosa`` (un-pinned in the
    reference's requirements.txt:3 and absent from this image) -- "parity unpinned"
    against librosa itself; the restatement follows librosa>=0.10's published algorithm
    and is pinned against the independent ``torch.istft`` / ``torch.stft`` in float64.

Every function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# 1. Sampler: schedule tables, respacing, noise layout, step, loop
# --------------------------------------------------------------------------------------


class Schedule:
    """float64 diffusion tables.  model/DiffSynthSampler.py:55-57,169-190 (tables) and
    :204-222 (respace)."""

    def __init__(self, timesteps: int, beta_start: float = 1e-4, beta_end: float = 0.02):
        self.timestep_map = list(range(timesteps))
        self._set(np.linspace(beta_start, beta_end, timesteps).astype(np.float64))
        self.respaced = False

    def _set(self, betas: np.ndarray) -> None:
        self.betas = betas
        self.num_timesteps = len(betas)
        ac = np.cumprod(1.0 - betas)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.concatenate([[1.0], ac[:-1]])
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)

    def respace(self, use_timesteps: Sequence[int]) -> None:
        assert not self.respaced, "This schedule has already been respaced!"
        keep = set(int(i) for i in use_timesteps)
        last, betas, tmap = 1.0, [], []
        for i, a in enumerate(self.alphas_cumprod):
            if i in keep:
                betas.append(1.0 - a / last)
                last = a
                tmap.append(i)
        self.timestep_map = tmap
        self._set(np.asarray(betas, dtype=np.float64))
        # the reference sets num_timesteps = len(use_timesteps) (:219); duplicates in
        # use_timesteps (possible with linspace->int) make that differ from len(betas)
        self.num_timesteps = len(use_timesteps)
        self.respaced = True


def noise_layout_repeat(train_noise: Tensor, batch: int, width: int, train_width: int = 64
                        ) -> Tuple[Tensor, List[int]]:
    """'repeat' noise strategy: re-lay the columns of a [*, C, H, train_width] draw to
    ``width`` columns.  model/DiffSynthSampler.py:97-167."""
    rel = int(train_width * 1.0 / 4)
    body_w = train_width - rel
    body = train_noise[:batch, :, :, :body_w]
    tail = train_noise[:batch, :, :, train_width - rel:]
    if width <= train_width:
        head_w = int((width - rel) / 2)
        tail_w = width - rel - head_w
        parts = [body[..., :head_w], body[..., body_w - tail_w:] if tail_w > 0 else body[..., :0], tail]
        if tail_w == 0:
            # python's x[-0:] is the whole axis (reference :127); keep that quirk
            parts[1] = body
    else:
        reps = (width - rel) // body_w
        extra = (width - rel) % body_w
        hw = int(body_w / 2)
        tw = body_w - hw
        mid0 = (body_w - extra) // 2
        parts = [body[..., :hw]] * reps + [body[..., mid0:mid0 + extra]] + [body[..., body_w - tw:]] * reps + [tail]
    pts = [0]
    for p in parts[:-1]:
        pts.append(pts[-1] + p.shape[3])
    return torch.cat(parts, dim=3), pts


def q_sample(sch: Schedule, x0: Tensor, t: int, noise: Tensor) -> Tensor:
    """model/DiffSynthSampler.py:271-294 (coefficients f64 -> fp32 at use, :17)."""
    a = torch.tensor(sch.sqrt_alphas_cumprod[t]).float()
    b = torch.tensor(sch.sqrt_one_minus_alphas_cumprod[t]).float()
    return a * x0 + b * noise


def ddim_coefficients(sch: Schedule, t: int, eta: float) -> Dict[str, float]:
    """The per-step scalars of model/DiffSynthSampler.py:323-343, evaluated the way the
    reference does: alphas gathered in f64, cast to fp32 (:17), then fp32 torch math."""
    at = torch.tensor(sch.alphas_cumprod[t]).float()
    ap = torch.tensor(sch.alphas_cumprod_prev[t]).float()
    sigma = eta * torch.sqrt((1 - ap) / (1 - at)) * torch.sqrt(1 - at / ap)
    return dict(sqrt_one_minus_at=float(torch.sqrt(1.0 - at)), sqrt_at=float(torch.sqrt(at)),
                sqrt_ap=float(torch.sqrt(ap)), dir_coef=float(torch.sqrt(1 - ap - sigma ** 2)),
                sigma=float(sigma))


def ddim_update(x: Tensor, eps_u: Optional[Tensor], eps_c: Tensor, cfg_scale: float, co: Dict[str, float],
                step_noise: Tensor) -> Tensor:
    """CFG combine + DDIM/DDPM update.  model/DiffSynthSampler.py:320,327,337,343."""
    eps = eps_c if eps_u is None else eps_u + cfg_scale * (eps_c - eps_u)
    f = lambda v: torch.tensor(v, dtype=torch.float32)
    x0 = (x - f(co["sqrt_one_minus_at"]) * eps) / f(co["sqrt_at"])
    return f(co["sqrt_ap"]) * x0 + f(co["dir_coef"]) * eps + f(co["sigma"]) * step_noise


def dynamic_masks(n_masks: int, shape, concat_points: Sequence[int], mask_flexivity: float = 0.8,
                  train_width: int = 64) -> List[Tensor]:
    """get_dynamic_masks, model/DiffSynthSampler.py:365-422: freeze masks (1 = keep the guide) that shrink over the
    first ``int(n_masks * mask_flexivity)`` blends towards the release area only; returned in pop() order."""
    rel = int(train_width / 4)
    assert shape[3] == concat_points[-1] + rel, "shape[3] != (concat_points[-1] + release_length)"
    seg = [concat_points[i + 1] - concat_points[i] for i in range(len(concat_points) - 1)]
    n_guided = int(n_masks * mask_flexivity)
    masks = []
    for i in range(n_guided):
        m = torch.zeros((shape[0], 1, shape[2], shape[3]))
        m[..., shape[3] - rel:] = 1.0
        for k, length in enumerate(seg):
            keep = int((n_guided - 1 - i) / (n_guided - 1) * length)
            if k == 0:
                m[..., :keep] = 1.0
            elif k == len(seg) - 1:
                if keep != 0:
                    m[..., shape[3] - keep - rel:] = 1.0
            else:
                s0 = concat_points[k] + int((length - keep) / 2)
                m[..., s0:s0 + keep] = 1.0
        masks.append(m)
    for _ in range(n_masks - n_guided):
        m = torch.zeros((shape[0], 1, shape[2], shape[3]))
        m[..., shape[3] - rel:] = 1.0
        masks.append(m)
    masks.reverse()
    return masks


def sample_loop(model, sch: Schedule, shape, cond: Tensor, uncond: Optional[Tensor], cfg_scale: float,
                noise_draws: Tensor, sampler: str = "ddim", guide: Optional[Tensor] = None,
                start_ratio: float = 1.0, end_ratio: float = 0.0, mask: Optional[Tensor] = None,
                inpaint: bool = False, train_width: int = 64, trace: Optional[list] = None,
                dynamic_mask_flexivity: Optional[float] = None) -> List[Tensor]:
    """p_sample_loop with host-fed noise.  model/DiffSynthSampler.py:425-517 (+ :297-363).

    ``noise_draws`` [1+steps, B, C, H, train_width]: draw 0 is the initial noise, draw k the
    k-th per-step draw in loop order (the reference draws one per step even when eta=0, :340).
    ``model(x, t_mapped[int64 B], cond) -> eps``.  Returns the list of latents (len steps+1).
    If ``trace`` is a list, per-step dicts (t, x, eps_u, eps_c, x_prev) are appended."""
    if sampler not in ("ddim", "ddpm"):
        raise NotImplementedError()
    eta = 0.0 if sampler == "ddim" else 1.0
    B, W = shape[0], shape[3]
    init, _ = noise_layout_repeat(noise_draws[0], B, W, train_width)
    assert tuple(init.shape) == tuple(shape), "initial_noise.shape != shape"
    start = int(sch.num_timesteps * start_ratio)
    end = int(sch.num_timesteps * end_ratio)
    assert start_ratio == 1.0 or guide is not None
    if guide is None:
        img = init
    else:
        guide, concat_points = noise_layout_repeat(guide, B, W, train_width)
        img = q_sample(sch, guide, start - 1, init) if start > 0 else guide
    # :483-487: one mask per blend, popped from the end; the caller's mask is ignored when the dynamic ones are used
    masks = dynamic_masks(start - end, shape, concat_points, dynamic_mask_flexivity, train_width) \
        if dynamic_mask_flexivity is not None else [mask] * (start - end)
    imgs = [img]
    k = 1
    for i in reversed(range(end, start)):
        t_mapped = torch.full((B,), sch.timestep_map[i], dtype=torch.long)
        if cfg_scale == 1.0:
            eps_u, eps_c = None, model(img, t_mapped, cond)
        else:
            u = uncond.unsqueeze(0).repeat(*([B] + [1] * uncond.dim()))      # :313-314 (a vector for text conditions, a scalar label for class conditions)
            out = model(torch.cat([img, img]), torch.cat([t_mapped, t_mapped]), torch.cat([u, cond]))
            eps_u, eps_c = out[:B], out[B:]
        z, _ = noise_layout_repeat(noise_draws[k], B, W, train_width)
        k += 1
        new = ddim_update(img, eps_u, eps_c, cfg_scale, ddim_coefficients(sch, i, eta), z)
        if trace is not None:
            trace.append(dict(t=i, x=img, eps_u=eps_u, eps_c=eps_c, x_prev=new))
        img = new
        if inpaint:
            if i > 0:
                mask = masks.pop()
                img = mask * q_sample(sch, guide, i - 1, init) + (1 - mask) * img
            else:
                img = mask * guide + (1 - mask) * img
        imgs.append(img)
    return imgs


# --------------------------------------------------------------------------------------
# 2. U-Net  (model/diffusion.py:187-258 over model/diffusion_components.py)
# --------------------------------------------------------------------------------------

def _gn(x: Tensor, sd: SD, p: str, groups: int = 1, eps: float = 1e-5) -> Tensor:
    return F.group_norm(x, groups, sd[p + "weight"], sd[p + "bias"], eps)


def time_embedding(sd: SD, t: Tensor, dim: int) -> Tensor:
    """SinusoidalPositionEmbeddings + time_mlp.  diffusion_components.py:42-56; diffusion.py:100-105."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half) * -k)
    arg = t[:, None] * freqs[None, :]          # int64 * fp32 -> fp32
    e = torch.cat([arg.sin(), arg.cos()], dim=-1)
    e = F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    return F.linear(F.gelu(e), sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def convnext_block(sd: SD, p: str, x: Tensor, temb: Optional[Tensor]) -> Tensor:
    """ConvNextBlock.forward.  diffusion_components.py:130-139 (ctor :110-128)."""
    h = F.conv2d(x, sd[p + "ds_conv.weight"], sd[p + "ds_conv.bias"], padding=3, groups=x.shape[1])
    if temb is not None and (p + "mlp.1.weight") in sd:
        h = h + F.linear(F.gelu(temb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])[:, :, None, None]
    h = _gn(h, sd, p + "net.0.")
    h = F.gelu(F.conv2d(h, sd[p + "net.1.weight"], sd[p + "net.1.bias"], padding=1))
    h = _gn(h, sd, p + "net.3.")
    h = F.conv2d(h, sd[p + "net.4.weight"], sd[p + "net.4.bias"], padding=1)
    if (p + "res_conv.weight") in sd:
        return h + F.conv2d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return h + x


def resnet_block(sd: SD, p: str, x: Tensor, temb: Optional[Tensor], groups: int = 8) -> Tensor:
    """ResnetBlock.forward (use_convnext=False).  diffusion_components.py:59-104: Block = conv3x3 -> GroupNorm(groups) -> SiLU;
    the time embedding (SiLU -> Linear) is added between the two blocks."""
    def blk(h, q):
        h = F.conv2d(h, sd[q + "proj.weight"], sd[q + "proj.bias"], padding=1)
        return F.silu(F.group_norm(h, groups, sd[q + "norm.weight"], sd[q + "norm.bias"], eps=1e-5))
    h = blk(x, p + "block1.")
    if temb is not None and (p + "mlp.1.weight") in sd:
        h = F.linear(F.silu(temb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])[:, :, None, None] + h
    h = blk(h, p + "block2.")
    if (p + "res_conv.weight") in sd:
        return h + F.conv2d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return h + x


def _block(sd: SD, p: str, x: Tensor, temb: Optional[Tensor]) -> Tensor:
    """block_klass (diffusion.py:83-87), read off the state_dict."""
    if (p + "block1.proj.weight") in sd:
        return resnet_block(sd, p, x, temb)
    return convnext_block(sd, p, x, temb)


def linear_attention_add(sd: SD, p: str, x: Tensor, cemb: Optional[Tensor], heads: int = 4, dh: int = 32) -> Tensor:
    """Residual(PreNorm(LinearCrossAttentionAdd)).  diffusion_components.py:22-29,142-152,252-293."""
    B, C, H, W = x.shape
    n = H * W
    xn = _gn(x, sd, p + "fn.norm.")
    qkv = F.conv2d(xn, sd[p + "fn.fn.to_qkv.weight"]).reshape(B, 3, heads, dh, n)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    if cemb is not None:
        k = k + F.linear(cemb, sd[p + "fn.fn.label_key.weight"], sd[p + "fn.fn.label_key.bias"]).view(B, heads, dh, 1)
        q = q + F.linear(cemb, sd[p + "fn.fn.label_query.weight"], sd[p + "fn.fn.label_query.bias"]).view(B, heads, dh, 1)
    q = q.softmax(dim=-2) * dh ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, heads * dh, H, W)
    out = F.conv2d(out, sd[p + "fn.fn.to_out.0.weight"], sd[p + "fn.fn.to_out.0.bias"])
    return _gn(out, sd, p + "fn.fn.to_out.1.") + x


def linear_attention_cat(sd: SD, p: str, x: Tensor, cemb: Optional[Tensor], heads: int = 4, dh: int = 32) -> Tensor:
    """Residual(PreNorm(LinearCrossAttention)), attn_type="linear_cat".  diffusion_components.py:22-29,142-152,171-207:
    the condition contributes one extra key / value token instead of being added to q and k."""
    B, C, H, W = x.shape
    n = H * W
    xn = _gn(x, sd, p + "fn.norm.")
    qkv = F.conv2d(xn, sd[p + "fn.fn.to_qkv.weight"]).reshape(B, 3, heads, dh, n)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    if cemb is not None:
        lk = F.linear(cemb, sd[p + "fn.fn.label_key.weight"], sd[p + "fn.fn.label_key.bias"]).view(B, heads, dh, 1)
        lv = F.linear(cemb, sd[p + "fn.fn.label_value.weight"], sd[p + "fn.fn.label_value.bias"]).view(B, heads, dh, 1)
        k, v = torch.cat([k, lk], dim=-1), torch.cat([v, lv], dim=-1)
    q = q.softmax(dim=-2)
    k = k.softmax(dim=-1)
    q = q * dh ** -0.5
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, heads * dh, H, W)
    out = F.conv2d(out, sd[p + "fn.fn.to_out.0.weight"], sd[p + "fn.fn.to_out.0.bias"])
    return _gn(out, sd, p + "fn.fn.to_out.1.") + x


def _attention(sd: SD, p: str, x: Tensor, cemb: Optional[Tensor]) -> Tensor:
    """The variant is read off the state_dict: label_value exists only in LinearCrossAttention ("linear_cat")."""
    if (p + "fn.fn.label_value.weight") in sd:
        return linear_attention_cat(sd, p, x, cemb)
    return linear_attention_add(sd, p, x, cemb)


def pad_and_concat(enc: Tensor, dec: Tensor) -> Tensor:
    """diffusion_components.py:210-249: zero-pad ``dec`` to ``enc``'s H,W, cat (enc first)."""
    dh, dw = enc.shape[2] - dec.shape[2], enc.shape[3] - dec.shape[3]
    if dh or dw:
        dec = F.pad(dec, (dw // 2, dw - dw // 2, dh // 2, dh - dh // 2))
    return torch.cat([enc, dec], dim=1)


def unet_forward(sd: SD, x: Tensor, t: Tensor, cond: Optional[Tensor], taps: Optional[dict] = None) -> Tensor:
    """ConditionedUnet.forward, ConvNeXt or ResNet blocks with linear_add or linear_cat attention.  model/diffusion.py:187-258.
    The architecture is read off the state_dict keys.  ``taps`` (optional dict) receives
    named intermediates for per-layer parity tests."""
    n_stage = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("downs."))
    n_midl = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("mid_left."))
    dim0 = sd["init_conv.weight"].shape[0]

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    cemb = None
    if cond is not None:
        if "label_embedding.embedding.bias" in sd:      # condition_type "natural_language_prompt": nn.Linear (diffusion_components.py:163)
            cemb = F.linear(cond, sd["label_embedding.embedding.weight"], sd["label_embedding.embedding.bias"])
        else:                                           # "instrument_family": nn.Embedding over integer labels (:161)
            cemb = sd["label_embedding.embedding.weight"][cond.long()]
    hs = []
    x = tap("init_conv", F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3))
    hs.append(x)
    temb = tap("time_emb", time_embedding(sd, t, dim0)) if "time_mlp.1.weight" in sd else None      # with_time_emb=False: diffusion.py:107-109,211
    for i in range(n_stage):
        p = f"downs.{i}."
        x = tap(p + "0", _block(sd, p + "0.", x, temb))
        x = tap(p + "1", _attention(sd, p + "1.", x, cemb)); hs.append(x)
        x = tap(p + "2", _block(sd, p + "2.", x, temb))
        x = tap(p + "3", _attention(sd, p + "3.", x, cemb)); hs.append(x)
        x = tap(p + "4", F.conv2d(x, sd[p + "4.weight"], sd[p + "4.bias"], stride=2, padding=1)); hs.append(x)
    for j in range(n_midl):
        x = tap(f"mid_left.{j}", _block(sd, f"mid_left.{j}.", x, temb)); hs.append(x)
    x = tap("mid_mid.0", _block(sd, "mid_mid.0.", x, temb))
    x = tap("mid_mid.1", _attention(sd, "mid_mid.1.", x, cemb))
    x = tap("mid_mid.2", _block(sd, "mid_mid.2.", x, temb))
    for j in range(n_midl):
        x = tap(f"mid_right.{j}", _block(sd, f"mid_right.{j}.", pad_and_concat(hs.pop(), x), temb))
    for i in range(n_stage):
        p = f"ups.{i}."
        x = tap(p + "0", _block(sd, p + "0.", pad_and_concat(hs.pop(), x), temb))
        x = tap(p + "1", _attention(sd, p + "1.", x, cemb))
        x = tap(p + "2", F.conv_transpose2d(x, sd[p + "2.weight"], sd[p + "2.bias"], stride=2, padding=1))
        x = tap(p + "3", _block(sd, p + "3.", pad_and_concat(hs.pop(), x), temb))
        x = tap(p + "4", _attention(sd, p + "4.", x, cemb))
        x = tap(p + "5", _block(sd, p + "5.", pad_and_concat(hs.pop(), x), temb))
        x = tap(p + "6", _attention(sd, p + "6.", x, cemb))
    x = tap("final_conv.0", _block(sd, "final_conv.0.", pad_and_concat(hs.pop(), x), None))
    return F.conv2d(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"], padding=1)


# --------------------------------------------------------------------------------------
# 3. VQGAN: quantiser, decoder, encoder  (model/VQGAN.py)
# --------------------------------------------------------------------------------------

def _fma32(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """fp32 fused multiply-add emulated through float64 (the product of two fp32 is exact
    in f64; the single extra rounding f64->f32 differs from a true FMA with probability
    ~2^-29 per operation)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def vq_distances(flat: np.ndarray, codebook: np.ndarray) -> np.ndarray:
    """fp32 distances exactly as the CUDA kernel and (bitwise, on the CPUs tried) torch's
    CPU path evaluate model/VQGAN.py:107-109:
        (sum_d x_d^2 [N,1]  +  sum_d e_d^2 [K])  -  2 * (x @ E^T)
    with the K=4 reductions as ascending-d sums and the matmul as an ascending-d FMA chain."""
    x = flat.astype(np.float32)
    e = codebook.astype(np.float32)
    xs = x * x
    es = e * e
    sx = xs[:, 0]
    se = es[:, 0]
    for d in range(1, x.shape[1]):
        sx = sx + xs[:, d]
        se = se + es[:, d]
    dot = (x[:, 0:1] * e[None, :, 0]).astype(np.float32)
    for d in range(1, x.shape[1]):
        dot = _fma32(x[:, d:d + 1], e[None, :, d], dot)
    return (sx[:, None] + se[None, :]) - np.float32(2.0) * dot


def vq_quantize(latents: Tensor, codebook: Tensor, chunk: int = 2048) -> Tuple[Tensor, Tensor]:
    """VectorQuantizerEMA.forward, eval branch.  model/VQGAN.py:98-146.
    Returns (quantized NCHW fp32 = x + (q - x) as the straight-through line :134 leaves it,
    indices int64 [B*H*W] in NHWC-flattened order :100-104).  Chunked: the reference's
    [N,8192] fp32 matrices do not fit at B=64."""
    B, C, H, W = latents.shape
    flat = latents.permute(0, 2, 3, 1).contiguous().view(-1, C).numpy()
    cb = codebook.numpy().astype(np.float32)
    idx = np.empty(flat.shape[0], dtype=np.int64)
    for s in range(0, flat.shape[0], chunk):
        idx[s:s + chunk] = np.argmin(vq_distances(flat[s:s + chunk], cb), axis=1)   # first minimum
    q = cb[idx]
    out = flat + (q - flat)
    out = torch.from_numpy(out).view(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    return out, torch.from_numpy(idx)


def _swish_or_relu(x: Tensor, act: str) -> Tensor:
    """model/VQGAN.py:20-27."""
    return F.relu(x) if act == "relu" else x * torch.sigmoid(x)


def vq_resblock(sd: SD, p: str, x: Tensor, groups: int, act: str) -> Tensor:
    """VQGAN.ResnetBlock.forward with double_conv=False, temb=None.  model/VQGAN.py:223-244."""
    h = F.group_norm(x, groups, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    h = F.conv2d(_swish_or_relu(h, act), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    if (p + "nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def vq_linear_attention(sd: SD, p: str, x: Tensor, dh: int = 32) -> Tensor:
    """VQGAN.LinearAttention.forward (heads=1; q neither soft-maxed nor scaled).  model/VQGAN.py:261-272."""
    B, C, H, W = x.shape
    qkv = F.conv2d(x, sd[p + "to_qkv.weight"]).reshape(B, 3, dh, H * W)
    q, k, v = qkv[:, 0], qkv[:, 1].softmax(dim=-1), qkv[:, 2]
    ctx = torch.einsum("bdn,ben->bde", k, v)
    out = torch.einsum("bde,bdn->ben", ctx, q).reshape(B, dh, H, W)
    out = F.conv2d(out, sd[p + "to_out.weight"], sd[p + "to_out.bias"])
    if (p + "nin_shortcut.weight") in sd:
        out = out + F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return out


def _vq_stack(sd: SD, prefix: str, plan, x: Tensor, groups: int, act: str, taps: Optional[dict]) -> Tensor:
    for idx, kind, cin, cout in plan:
        p = f"{prefix}{idx}."
        if kind == "down":
            x = F.conv2d(x, sd[p + "_conv2d.weight"], sd[p + "_conv2d.bias"], stride=2, padding=1)
        elif kind == "up":
            x = F.conv_transpose2d(x, sd[p + "_conv2d.weight"], sd[p + "_conv2d.bias"], stride=2, padding=1)
        elif kind == "res":
            x = vq_resblock(sd, p, x, groups, act)
        elif kind == "attn":
            x = vq_linear_attention(sd, p, x)
        elif kind == "norm":
            x = F.group_norm(x, groups, sd[p + "weight"], sd[p + "bias"], 1e-6)
        elif kind == "relu":
            x = F.relu(x)
        elif kind == "conv1x1":
            x = F.conv2d(x, sd[p + "weight"], sd[p + "bias"])
        elif kind == "conv1x1_nobias":
            x = F.conv2d(x, sd[p + "weight"])
        if taps is not None:
            taps[f"{prefix}{idx}"] = x
    return x


def vqgan_decode(sd: SD, plan, z: Tensor, groups: int = 16, act: str = "swish", taps: Optional[dict] = None) -> Tensor:
    """Decoder.forward: layer stack, then softplus on ch0 and tanh on ch1/ch2.  model/VQGAN.py:390-400."""
    x = _vq_stack(sd, "_decoder._layers.", plan, z, groups, act, taps)
    return torch.stack([F.softplus(x[:, 0]), torch.tanh(x[:, 1]), torch.tanh(x[:, 2])], dim=1)


def vqgan_encode(sd: SD, plan, spec: Tensor, groups: int = 16, taps: Optional[dict] = None) -> Tensor:
    """Encoder.forward.  model/VQGAN.py:323-326.  The encoder's ResnetBlocks run swish
    whatever the config says: VQGAN.__init__ passes the literal string "act_type" (:441)."""
    return _vq_stack(sd, "_encoder._layers.", plan, spec, groups, "act_type", taps)


# --------------------------------------------------------------------------------------
# 4. STFT+ codec and (i)STFT   (tools.py; librosa call sites in webUI/.../utils.py:241)
# --------------------------------------------------------------------------------------

def decode_stft(enc: np.ndarray) -> np.ndarray:
    """tools.py:334-345."""
    mag = np.expm1(enc[0])
    ph = np.arctan2(enc[2], enc[1])
    return mag * (np.cos(ph) + 1j * np.sin(ph))


def encode_stft(D: np.ndarray) -> np.ndarray:
    """tools.py:320-331."""
    ph = np.angle(D)
    return np.stack([np.log1p(np.abs(D)), np.cos(ph), np.sin(ph)], axis=0)


def depad_stft(Dp: np.ndarray) -> np.ndarray:
    """tools.py:185-191 (the zero DC row is float64, so the result is complex128)."""
    return np.concatenate([np.zeros((1, Dp.shape[1])), Dp], axis=0)


def pad_stft(D: np.ndarray, time_resolution: Optional[int] = 256) -> np.ndarray:
    """tools.py:170-182."""
    D = D[1:, :]
    if time_resolution is None or time_resolution - D.shape[1] <= 0:
        return D
    return np.pad(D, ((0, 0), (0, time_resolution - D.shape[1])), "constant")


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True), the window librosa uses."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def istft(D: np.ndarray, hop: int = 256, win: int = 1024) -> np.ndarray:
    """librosa.istft(D, hop_length=hop, win_length=win) as called at
    webUI/natural_language_guided_4/utils.py:241 (center=True, window='hann', length=None):
    n_fft = 2*(rows-1); per-frame irfft * window; overlap-add; divide by the overlap-added
    squared window where it exceeds ``tiny``; drop n_fft//2 samples at both ends."""
    n_fft = 2 * (D.shape[0] - 1)
    assert win == n_fft
    T = D.shape[1]
    w = hann_periodic(win)
    frames = np.fft.irfft(D.astype(np.complex128), n=n_fft, axis=0) * w[:, None]     # [n_fft, T]
    total = n_fft + hop * (T - 1)
    y = np.zeros(total)
    wss = np.zeros(total)
    for t in range(T):
        y[t * hop:t * hop + n_fft] += frames[:, t]
        wss[t * hop:t * hop + n_fft] += w * w
    ok = wss > np.finfo(np.float64).tiny
    y[ok] /= wss[ok]
    return y[n_fft // 2: total - n_fft // 2]


def stft(y: np.ndarray, n_fft: int = 1024, hop: int = 256, pad_mode: str = "constant") -> np.ndarray:
    """librosa.stft(y, n_fft=1024, hop_length=256, win_length=1024) as called at
    webUI/natural_language_guided_4/sound2sound_with_text.py:85 (center=True; librosa>=0.10
    pads with zeros, older releases reflect -- ``pad_mode`` exposes the choice)."""
    yp = np.pad(np.asarray(y, dtype=np.float64), n_fft // 2, mode=pad_mode)
    T = 1 + (len(yp) - n_fft) // hop
    w = hann_periodic(n_fft)
    frames = np.stack([yp[t * hop:t * hop + n_fft] * w for t in range(T)], axis=1)
    return np.fft.rfft(frames, axis=0)


def spectrogram_to_waveform(spec3: np.ndarray) -> np.ndarray:
    """decode_stft -> depad_STFT -> istft, one sample [3, 512, T].  utils.py:229-241."""
    return istft(depad_stft(decode_stft(spec3)))


def waveform_to_spectrogram(y: np.ndarray, time_resolution: int = 256) -> np.ndarray:
    """stft -> pad_STFT -> encode_stft.  sound2sound_with_text.py:85-94."""
    return encode_stft(pad_stft(stft(y), time_resolution)).astype(np.float32)


# --------------------------------------------------------------------------------------
# 5. Image products of the decode glue   (webUI/natural_language_guided_4/utils.py:8-128)
# --------------------------------------------------------------------------------------

def _np_log10(x: np.ndarray) -> np.ndarray:
    """tools.py:11-15."""
    return np.log(x + 1e-16) / np.log(10)


def np_power_to_db(S: np.ndarray, amin: float = 1e-16, top_db: float = 80.0) -> np.ndarray:
    """tools.py:41-50."""
    ref = S.max()
    log_spec = 10.0 * _np_log10(np.maximum(amin, S))
    log_spec -= 10.0 * _np_log10(np.maximum(amin, ref))
    return np.maximum(log_spec, log_spec.max() - top_db)


def _to_u8(x: np.ndarray) -> np.ndarray:
    """``.astype(np.uint8)`` of a float array as numpy does it on x86-64 (truncate, keep the low byte)."""
    with np.errstate(invalid="ignore"):
        return x.astype(np.uint8)


def spectrogram_image(spc: np.ndarray) -> np.ndarray:
    """spectrogram_to_Gradio_image, utils.py:8-50: |D| [F, T] -> uint8 [F, T, 3]."""
    nf, T = spc.shape[-2], spc.shape[-1]
    flipped = np.flipud(np_power_to_db(np.abs(np.reshape(spc, (nf, T)))))
    img = np.ones((nf, T, 3)) * -80.0
    img[:, :, 0] = flipped
    img[:, :, 1] = flipped
    img[:, :, 2] = -60.0
    return _to_u8(255.0 * ((img + 80.0) / 80.0))


def phase_image(phase: np.ndarray) -> np.ndarray:
    """phase_to_Gradio_image, utils.py:53-91: angle(D) [F, T] -> uint8 [F, T, 3]."""
    nf, T = phase.shape[-2], phase.shape[-1]
    flipped = (np.flipud(np.reshape(phase, (nf, T))) + 1.0) / 2.0
    img = np.zeros((nf, T, 3))
    img[:, :, 0] = flipped
    img[:, :, 1] = flipped
    img[:, :, 2] = 0.2
    return _to_u8(255.0 * img)


def latent_image(latent: np.ndarray) -> np.ndarray:
    """latent_representation_to_Gradio_image, utils.py:94-128: [4, H, W] float32 -> uint8 [8H, 8W, 4]
    (works on a copy; the reference normalises its argument in place)."""
    image = np.array(latent, dtype=np.float32, copy=True)
    for c in range(4):
        ch = image[c]
        with np.errstate(invalid="ignore", divide="ignore"):
            image[c] = (ch - ch.min()) / (ch.max() - ch.min()) * 255
    big = np.repeat(np.repeat(np.transpose(image, (1, 2, 0)), 8, axis=0), 8, axis=1)
    return _to_u8(np.flipud(big))


def decode_products(spec3: np.ndarray):
    """The loop body of encodeBatch2GradioOutput_STFT / InputBatch2Encode_STFT for one sample [3, 512, T]
    (utils.py:229-241): -> (spectrogram image, phase image, waveform)."""
    D = depad_stft(decode_stft(spec3))
    return spectrogram_image(np.abs(D)), phase_image(np.angle(D)), istft(D)


# --------------------------------------------------------------------------------------
# 6. Griffin-Lim   (librosa.griffinlim as called at tools.py:75,214,222: n_iter 32/50, hop 256, win 1024)
# --------------------------------------------------------------------------------------

def griffinlim(S: np.ndarray, init_phase: np.ndarray, n_iter: int = 32, momentum: float = 0.99) -> np.ndarray:
    """Third-party arithmetic (librosa, unpinned, absent here: parity unpinned).  Restatement of the published
    fast Griffin-Lim loop (Perraudin et al. 2013) as librosa >= 0.7 implements it: angles_0 = exp(i * init_phase)
    (librosa draws init_phase = 2*pi*U[0,1) from an unseeded generator); each iteration
    rebuilt = stft(istft(S * angles)); angles = rebuilt - momentum/(1+momentum) * tprev (from the second iteration on);
    angles /= |angles| + tiny(float32); tprev = rebuilt; result = istft(S * angles).  S is the [513, T] magnitude."""
    angles = np.exp(1j * init_phase) * S
    tprev = None
    eps = np.finfo(np.float32).tiny
    for _ in range(n_iter):
        rebuilt = stft(istft(angles), pad_mode="constant")
        angles = rebuilt.copy()
        if tprev is not None:
            angles = angles - (momentum / (1 + momentum)) * tprev
        angles = angles / (np.abs(angles) + eps)
        angles = angles * S
        tprev = rebuilt
    return istft(angles)


# ---------------------------------------------------------------------------------------------------------------------
# Text-conditioning front end.  The text tower is THIRD-PARTY code: transformers' ClapModel
# (requirements.txt:5, unpinned; app.py:44 loads "laion/clap-htsat-unfused") -- models/clap/modeling_clap.py: ClapTextEmbeddings
# (word + token-type + position embeddings with RoBERTa's padding-aware position ids, LayerNorm), 12 x ClapTextLayer (self
# attention, post-LayerNorm residual blocks, GELU(erf) feed-forward), ClapTextPooler (tanh dense on token 0),
# ClapProjectionLayer (linear, ReLU, linear) and F.normalize in ClapModel.get_text_features.  The restatement is pinned against
# the installed transformers (5.5) ClapModel in tests/test_oracle_text.py.  ProjectionHead is the reference's own code
# (model/multimodal_model.py:14-47), applied by multi_modal_model.get_text_features (:114-116).
# ---------------------------------------------------------------------------------------------------------------------
def clap_text_features(sd: SD, input_ids: Tensor, attention_mask: Tensor, heads: int = 12, pad_idx: int = 1, eps: float = 1e-12,
                       num_projection_layers: int = 2, taps: Optional[dict] = None) -> Tensor:
    t = "text_encoder.text_model."
    mask = input_ids.ne(pad_idx).long()
    pos_ids = torch.cumsum(mask, dim=1) * mask + pad_idx                                   # create_position_ids_from_input_ids
    x = sd[t + "embeddings.word_embeddings.weight"][input_ids] + sd[t + "embeddings.token_type_embeddings.weight"][0] \
        + sd[t + "embeddings.position_embeddings.weight"][pos_ids]
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd[t + "embeddings.LayerNorm.weight"], sd[t + "embeddings.LayerNorm.bias"], eps)
    B, L, _ = x.shape
    dh = D // heads
    bias = (1.0 - attention_mask[:, None, None, :].to(x.dtype)) * torch.finfo(x.dtype).min  # additive key mask
    i = 0
    while f"{t}encoder.layer.{i}.attention.self.query.weight" in sd:
        p = f"{t}encoder.layer.{i}."
        q, k, v = (F.linear(x, sd[p + f"attention.self.{n}.weight"], sd[p + f"attention.self.{n}.bias"]).view(B, L, heads, dh).transpose(1, 2)
                   for n in ("query", "key", "value"))
        a = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5 + bias, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, L, D)
        x = F.layer_norm(F.linear(a, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]) + x, (D,),
                         sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], eps)
        f = F.gelu(F.linear(x, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        x = F.layer_norm(F.linear(f, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]) + x, (D,),
                         sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)
        i += 1
    if taps is not None:
        taps["last_hidden_state"] = x
    pooled = torch.tanh(F.linear(x[:, 0], sd[t + "pooler.dense.weight"], sd[t + "pooler.dense.bias"]))
    y = F.linear(F.relu(F.linear(pooled, sd["text_encoder.text_projection.linear1.weight"], sd["text_encoder.text_projection.linear1.bias"])),
                 sd["text_encoder.text_projection.linear2.weight"], sd["text_encoder.text_projection.linear2.bias"])
    y = F.normalize(y, dim=-1)
    if taps is not None:
        taps["clap_text_features"] = y
    for j in range(num_projection_layers):                                                 # ProjectionHead, multimodal_model.py:25-32,44-47
        p = f"text_projection.layers.{j}."
        projected = F.linear(y, sd[p + "projection.weight"], sd[p + "projection.bias"])
        y = F.linear(F.gelu(projected), sd[p + "fc.weight"], sd[p + "fc.bias"]) + projected
        y = F.layer_norm(y, (y.shape[-1],), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
    return y

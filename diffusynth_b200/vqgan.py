"""VQGAN pieces of the sampling path -- B200 drop-ins for model/VQGAN.py:

  VectorQuantizerEMA.forward (eval)  :98-146   -> (quantized, loss, (perplexity, None, None))
  Decoder.forward                    :390-400  -> [B,3,4H,4W] fp32 (softplus / tanh / tanh heads)
  Encoder.forward                    :323-326  -> [B,4,H/4,W/4] fp32
  VQGAN (container with _encoder/_vq_vae/_decoder, load_state_dict with the reference's keys)  :432-458

Convolutions (1x1, 3x3, 4x4 stride-2, transposed 4x4 stride-2) run on the tcgen05 implicit-GEMM kernel;
GroupNorm(16)+swish/ReLU, the linear attention core and the heads are bandwidth-bound kernels.  Activations are
bf16 NHWC with channel counts padded to a multiple of 32 (80 -> 96), the padding kept at zero."""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib, engine, ops, weights as W
from ._lib import check
from .ops import PackedConv, conv_args, pack_conv_down, pack_conv_s1, pack_conv_up, run_conv

DH = W.VQ_ATTN_DIM
GN_CHUNKS = 64


def _cp(c: int) -> int:
    return (c + 31) // 32 * 32


class VectorQuantizerEMA:
    def __init__(self, num_embeddings, embedding_dim, commitment_cost=0.25, decay=0.99, epsilon=1e-5, device=None):
        if embedding_dim != 4:
            raise NotImplementedError("embedding_dim must be 4 (deployed VQGAN, app.py:32)")
        self._num_embeddings, self._embedding_dim, self._commitment_cost = num_embeddings, embedding_dim, commitment_cost
        self.device = torch.device(device if device is not None else "cuda")
        self.codebook: Optional[torch.Tensor] = None
        self.training = False
        self.last_indices: Optional[torch.Tensor] = None
        # loss / perplexity of the eval branch (:130,:136-142) are computed like the reference does unless this is switched
        # off; every caller on the sampling path discards them (text2sound.py:128), so TextToTimbre asks for None instead
        self.compute_aux = True

    def load_codebook(self, weight: torch.Tensor):
        assert tuple(weight.shape) == (self._num_embeddings, self._embedding_dim)
        self.codebook = weight.detach().float().contiguous().to(self.device)
        return self

    @torch.no_grad()
    def forward(self, inputs, return_indices: bool = False):
        """inputs [B,4,H,W] fp32 -> (quantized [B,4,H,W], loss, (perplexity, None, None)).
        ``loss`` and ``perplexity`` are what the reference computes in eval mode (:130,:136-137); its callers
        discard them (text2sound.py:128), here they are derived from the indices / outputs with torch reductions."""
        x = inputs.to(self.device, torch.float32).contiguous()
        B, Cc, H, Wd = x.shape
        assert Cc == self._embedding_dim
        out = torch.empty_like(x)
        idx = torch.empty((B * H * Wd,), dtype=torch.long, device=self.device)
        self.quantize_into(x, out, idx)
        self.last_indices = idx
        if not self.compute_aux:
            return out, None, (None, None, idx if return_indices else None)
        q = self.codebook[idx].view(B, H, Wd, Cc).permute(0, 3, 1, 2)
        loss = self._commitment_cost * torch.mean((q - x) ** 2)
        probs = torch.bincount(idx, minlength=self._num_embeddings).float() / idx.numel()
        perplexity = torch.exp(-torch.sum(probs * torch.log(probs + 1e-10)))
        if return_indices:
            return out, loss, (perplexity, None, idx)
        return out, loss, (perplexity, None, None)

    __call__ = forward

    def quantize_into(self, x: torch.Tensor, out: torch.Tensor, idx: torch.Tensor) -> None:
        """The quantiser kernel alone on caller-owned buffers (capturable in a CUDA graph): x, out fp32 [B,4,H,W], idx int64 [B*H*W]."""
        B, Cc, H, Wd = x.shape
        check(_lib.load().ds_vq_quantize(x.data_ptr(), self.codebook.data_ptr(), self._num_embeddings, out.data_ptr(), idx.data_ptr(),
                                         B, H * Wd, ops._stream()), "ds_vq_quantize")

    def eval(self):
        return self


class _Stack:
    """Encoder or Decoder layer stack executed through the C ABI (static plan per input shape)."""

    def __init__(self, sd, prefix: str, plan, cfg, device, is_decoder: bool):
        self.prefix, self.plan_spec, self.cfg, self.device, self.is_decoder = prefix, plan, cfg, device, is_decoder
        self.G = cfg["num_groups"]
        # the encoder's ResnetBlocks get the literal string "act_type" (VQGAN.py:441) -> swish; decoder: cfg act_type
        self.res_act = 2 if (not is_decoder or cfg["act_type"] != "relu") else 1
        self.layers: Dict[int, dict] = {}
        for idx, kind, cin, cout in plan:
            p = f"{prefix}{idx}."
            L: dict = dict(kind=kind, cin=cin, cout=cout)
            if kind == "down":
                L["conv"] = pack_conv_down(sd[p + "_conv2d.weight"], sd[p + "_conv2d.bias"].float(), cin_pad=_cp(cin))
            elif kind == "up":
                L["conv"] = pack_conv_up(sd[p + "_conv2d.weight"], sd[p + "_conv2d.bias"].float(), cin_pad=_cp(cin))
            elif kind == "res":
                L["gamma"], L["beta"] = sd[p + "norm1.weight"].float().contiguous(), sd[p + "norm1.bias"].float().contiguous()
                L["conv"] = pack_conv_s1(sd[p + "conv1.weight"], sd[p + "conv1.bias"].float(), cin_pad=_cp(cin))
                if (p + "nin_shortcut.weight") in sd:
                    L["short"] = pack_conv_s1(sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"].float(), cin_pad=_cp(cin))
            elif kind == "attn":
                L["qkv"] = pack_conv_s1(sd[p + "to_qkv.weight"], None, cin_pad=_cp(cin))
                L["wout"] = sd[p + "to_out.weight"].float().reshape(cin, DH).contiguous()
                cout_pad = ops.pad16(cin)
                e2 = torch.zeros(1, cout_pad)
                e2[0, :cin] = sd[p + "to_out.bias"].float()
                L["out"] = PackedConv(weight=torch.zeros(1, dtype=ops.ACT), e2=e2, e1=None, taps=[[(0, 0, 0)]], cin=DH, cout=cin,
                                      cout_pad=cout_pad, ncls=1, kind="s1")
                if (p + "nin_shortcut.weight") in sd:
                    L["short"] = pack_conv_s1(sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"].float(), cin_pad=_cp(cin))
            elif kind == "norm":
                L["gamma"], L["beta"] = sd[p + "weight"].float().contiguous(), sd[p + "bias"].float().contiguous()
            elif kind in ("conv1x1", "conv1x1_nobias"):
                L["conv"] = pack_conv_s1(sd[p + "weight"], sd[p + "bias"].float() if kind == "conv1x1" else None, cin_pad=_cp(cin))
            for k, v in L.items():
                if isinstance(v, PackedConv):
                    v.to(device)
                elif isinstance(v, torch.Tensor):
                    L[k] = v.to(device)
            self.layers[idx] = L
        self._plans: Dict[Tuple, "_StackPlan"] = {}

    def plan(self, B, H, Wd):
        key = (B, H, Wd)
        if key not in self._plans:
            self._plans[key] = _StackPlan(self, B, H, Wd)
        return self._plans[key]


class _StackPlan:
    def __init__(self, st: _Stack, B: int, H: int, Wd: int):
        dev = st.device
        lib = _lib.load()
        stream = ops._stream
        self.ops: List[Tuple[str, Callable[[], None]]] = []
        self.keep: list = []
        self.named: Dict[str, Tuple[torch.Tensor, int]] = {}
        f32 = dict(dtype=torch.float32, device=dev)
        first_c = st.plan_spec[0][2]
        self.inp = torch.zeros((B, first_c, H, Wd), **f32)

        def act(h, w, c):      # zero-initialised: padded channels must stay zero
            return torch.zeros((B, h, w, _cp(c)), dtype=ops.ACT, device=dev)

        self.meta: Dict[str, dict] = {}       # per op: kernel family + algorithmic FLOPs / HBM bytes (bench.py's per-kernel rooflines)

        def add(name, fn, family="misc", flops=0.0, bytes=0.0):
            self.ops.append((name, fn))
            self.meta[name] = dict(family=family, flops=float(flops), bytes=float(bytes))

        def conv(name, pc, src, h, w, **kw):
            a, stt, keep = conv_args(pc, src, None, B, h, w, wide_tiles=st.cfg.get("batch_invariant", False), **kw)
            self.keep += keep + [a]
            fam, fl, by = ops.conv_cost(a, pc)
            add(name, lambda a=a: run_conv(a), "vqgan_" + fam, fl, by)
            return stt

        def gn_act(name, x, c, h, w, gamma, beta, actv):
            part = torch.empty((B, st.G, GN_CHUNKS, 2), **f32)
            out = act(h, w, c)
            cp = x.shape[-1]
            add(name + ".stats", lambda: check(lib.ds_group_stats(x.data_ptr(), part.data_ptr(), B, c, cp, st.G, h * w, GN_CHUNKS, stream()), "group_stats"),
                "group_stats", 0.0, B * h * w * cp * 2.0)
            add(name + ".apply", lambda: check(lib.ds_gn_act(x.data_ptr(), out.data_ptr(), part.data_ptr(), GN_CHUNKS, gamma.data_ptr(),
                                                            beta.data_ptr(), B, c, cp, st.G, h * w, 1e-6, actv, stream()), "gn_act"),
                "gn_act", 0.0, 2.0 * B * h * w * cp * 2.0)
            self.keep.append(part)
            return out

        h, w = H, Wd
        x = act(h, w, first_c)
        add("to_nhwc", lambda x=x: check(lib.ds_nchw_f32_to_nhwc_bf16(self.inp.data_ptr(), x.data_ptr(), B, first_c, x.shape[-1], H * Wd, stream()), "to_nhwc"),
            "nchw_to_nhwc", 0.0, B * H * Wd * (first_c * 4.0 + x.shape[-1] * 2.0))
        self.out_f32: Optional[torch.Tensor] = None
        pending_norm = None
        last_idx = st.plan_spec[-1][0]
        for idx, kind, cin, cout in st.plan_spec:
            L = st.layers[idx]
            name = f"{st.prefix}{idx}"
            if kind == "down":
                o = act(h // 2, w // 2, cout)
                conv(name, L["conv"], x, h, w, out=o)
                x, h, w = o, h // 2, w // 2
            elif kind == "up":
                o = act(2 * h, 2 * w, cout)
                conv(name, L["conv"], x, h, w, out=o)
                x, h, w = o, 2 * h, 2 * w
            elif kind == "res":
                t = gn_act(name + ".norm1", x, cin, h, w, L["gamma"], L["beta"], st.res_act)
                if cout % 16 == 0:
                    r = x
                    if "short" in L:
                        r = act(h, w, cout)
                        conv(name + ".short", L["short"], x, h, w, out=r)
                    o = act(h, w, cout)
                    conv(name, L["conv"], t, h, w, out=o, residual=r)
                    x = o
                else:
                    # final decoder block (80 -> 3): both branches as fp32 NCHW, summed inside the head kernel
                    assert idx == last_idx and st.is_decoder and "short" in L
                    a32 = torch.empty((B, cout, h, w), **f32)
                    b32 = torch.empty((B, cout, h, w), **f32)
                    conv(name, L["conv"], t, h, w, out_f32=a32)
                    conv(name + ".short", L["short"], x, h, w, out_f32=b32)
                    self.out_f32 = torch.empty((B, cout, h, w), **f32)
                    add("heads", lambda: check(lib.ds_decoder_head(a32.data_ptr(), b32.data_ptr(), self.out_f32.data_ptr(), B, h * w, stream()), "decoder_head"),
                        "decoder_head", 0.0, 3.0 * B * cout * h * w * 4.0)
                    x = None
            elif kind == "attn":
                npix = h * w
                qkv = act(h, w, 3 * DH)
                conv(name + ".to_qkv", L["qkv"], x, h, w, out=qkv)
                qp = act(h, w, DH)
                part = torch.empty((lib.ds_attn_part_floats(B, 1, npix),), **f32)
                M = torch.empty((B, L["out"].cout_pad, DH), dtype=ops.ACT, device=dev)
                add(name + ".ctx", lambda qkv=qkv, qp=qp, part=part, npix=npix: check(
                    lib.ds_attn_ctx_partial(qkv.data_ptr(), qp.data_ptr(), part.data_ptr(), B, 1, npix, 1, 1.0, stream()), "attn_ctx_partial"),
                    "attn_ctx", 2.0 * B * DH * DH * npix, B * npix * (3 * DH + DH) * 2.0)
                add(name + ".fin", lambda part=part, M=M, L=L, npix=npix, cin=cin: check(
                    lib.ds_attn_finalize(part.data_ptr(), L["wout"].data_ptr(), M.data_ptr(), B, 1, npix, cin, L["out"].cout_pad, stream()), "attn_finalize"),
                    "attn_finalize", 0.0, part.numel() * 4.0)
                r = None
                if "short" in L:
                    r = act(h, w, cin)
                    conv(name + ".short", L["short"], x, h, w, out=r)
                o = act(h, w, cin)
                conv(name + ".to_out", L["out"], qp, h, w, out=o, residual=r, weight_override=M, per_sample_weights=True)
                self.keep += [part, M]
                x = o
            elif kind == "norm":
                pending_norm = (L["gamma"], L["beta"], cin, name)
            elif kind == "relu":
                g, b_, c, nname = pending_norm
                x = gn_act(nname, x, c, h, w, g, b_, 1)       # Normalize + nn.ReLU fused into one apply pass
                pending_norm = None
            elif kind in ("conv1x1", "conv1x1_nobias"):
                if cout % 16 == 0:
                    o = act(h, w, cout)
                    conv(name, L["conv"], x, h, w, out=o)
                    x = o
                else:
                    self.out_f32 = torch.empty((B, cout, h, w), **f32)
                    conv(name, L["conv"], x, h, w, out_f32=self.out_f32)
                    x = None
            if x is not None and kind not in ("norm",):
                self.named[name] = (x, cout if kind not in ("relu",) else x.shape[-1])
        if self.out_f32 is None:
            raise NotImplementedError("layer stack must end in an fp32 output (decoder heads or encoder 1x1 conv)")

    def run(self):
        for _, fn in self.ops:
            fn()


class _StackModule:
    def __init__(self, stack: _Stack, eng: Optional["engine.VqganEngine"] = None):
        self._stack = stack
        self._engine = eng            # module-level C ABI handle (ds_vqgan_decode / ds_vqgan_encode); the stack plan stays for taps / timings
        self.use_engine = True
        self._probe = torch.zeros(1, device=stack.device)

    def parameters(self):
        return iter([self._probe])

    def eval(self):
        return self

    @torch.no_grad()
    def forward(self, x):
        x = x.to(self._stack.device, torch.float32)
        B, Cc, H, Wd = x.shape
        if self._engine is not None and self.use_engine:
            return (self._engine.decode if self._stack.is_decoder else self._engine.encode)(x.contiguous())
        pl = self._stack.plan(B, H, Wd)
        pl.inp.copy_(x)
        pl.run()
        return pl.out_f32.clone()

    __call__ = forward


class Decoder(_StackModule):
    """model/VQGAN.py:329-400."""


class Encoder(_StackModule):
    """model/VQGAN.py:275-326."""


class VQGAN:
    def __init__(self, in_channels, hidden_channels, embedding_dim, out_channels, block_depth=2, attn_pos=None, attn_with_skip=True,
                 norm_type="groupnorm", act_type="relu", num_embeddings=1024, commitment_cost=0.25, decay=0.99, num_groups=32, device=None, batch_invariant=False):
        if norm_type != "groupnorm":
            raise NotImplementedError("norm_type='batchnorm'")
        if not decay > 0.0:
            raise NotImplementedError("non-EMA VectorQuantizer (training-only variant)")
        self.cfg = dict(in_channels=in_channels, hidden_channels=list(hidden_channels), embedding_dim=embedding_dim, out_channels=out_channels,
                        block_depth=block_depth, attn_pos=list(attn_pos or []), attn_with_skip=attn_with_skip, norm_type=norm_type,
                        act_type=act_type, num_embeddings=num_embeddings, commitment_cost=commitment_cost, decay=decay, num_groups=num_groups,
                        batch_invariant=bool(batch_invariant))      # (see ConditionedUnet: tiling policy, not a reference argument)
        self.device = torch.device(device if device is not None else "cuda")
        self._vq_vae = VectorQuantizerEMA(num_embeddings, embedding_dim, commitment_cost, decay, device=self.device)
        self._encoder: Optional[Encoder] = None
        self._decoder: Optional[Decoder] = None
        self._sd = None
        self._engine = None

    def load_state_dict(self, state_dict, strict=True):
        spec = W.vqgan_param_spec(self.cfg)
        missing = [k for k, _ in spec if k not in state_dict]
        if strict and missing:
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:4]}...")
        for k, shp in spec:
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {shp}")
        sd = OrderedDict((k, state_dict[k].detach().float().cpu().contiguous()) for k, _ in spec if k in state_dict)
        self._sd = sd
        enc_plan, dec_plan = W.vqgan_layer_plan(self.cfg)
        self._vq_vae.load_codebook(sd["_vq_vae._embedding.weight"])
        self._engine = engine.VqganEngine(self.cfg, sd, self.device)
        self._encoder = Encoder(_Stack(sd, "_encoder._layers.", enc_plan, self.cfg, self.device, is_decoder=False), self._engine)
        self._decoder = Decoder(_Stack(sd, "_decoder._layers.", dec_plan, self.cfg, self.device, is_decoder=True), self._engine)
        return self

    def state_dict(self):
        return OrderedDict(self._sd)

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device:
            raise NotImplementedError("construct VQGAN with device=... (weights are packed for one device)")
        return self

    @torch.no_grad()
    def forward(self, x):
        z = self._encoder(x)
        quantized, vq_loss, (perplexity, _, _) = self._vq_vae(z)
        return vq_loss, self._decoder(quantized), perplexity

    __call__ = forward

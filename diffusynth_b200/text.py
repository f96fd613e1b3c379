"""Text-conditioning front end on the GPU: prompt token ids -> the [B, 512] condition vectors the sampler takes.

Drop-in for ``text_encoder.get_text_features(**tokenizer(prompts, padding=True, return_tensors="pt"))`` as the reference calls it
(webUI/natural_language_guided_4/text2sound.py:89-109), where ``text_encoder`` is the multi_modal_model of app.py:55-59:
``ClapModel.get_text_features`` (transformers: ClapTextModel = RoBERTa-base, ClapProjectionLayer, F.normalize) followed by
``ProjectionHead`` (model/multimodal_model.py:14-47,114-116).  With ``use_pretrained_CLAP`` (app.py:28,51-52) the head is absent:
construct with ``num_projection_layers=0``.

Unlike the reference (one prompt on the CPU, then ``.repeat``), a batch of DISTINCT prompts is encoded in one pass.  The
tokenizer itself is string processing on the host and stays with the caller (there is no vocabulary offline): inputs are the
tokenizer's ``input_ids`` / ``attention_mask``.  ``load_state_dict`` takes the multi_modal_model's keys
(``text_encoder.text_model...``, ``text_encoder.text_projection...``, ``text_projection.layers.N...``)."""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Optional, Tuple

import torch

from . import _lib, ops
from ._lib import check
from .ops import conv_args, pack_conv_s1, run_conv

CLAP_TEXT = dict(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 max_position_embeddings=514, pad_token_id=1, layer_norm_eps=1e-12, projection_dim=512)      # ClapTextConfig defaults


def text_param_spec(cfg: dict, num_projection_layers: int = 2, emb_dim: int = 512) -> List[Tuple[str, Tuple[int, ...]]]:
    D, I, P = cfg["hidden_size"], cfg["intermediate_size"], cfg["projection_dim"]
    t = "text_encoder.text_model."
    spec = [(t + "embeddings.word_embeddings.weight", (cfg["vocab_size"], D)), (t + "embeddings.token_type_embeddings.weight", (1, D)),
            (t + "embeddings.LayerNorm.weight", (D,)), (t + "embeddings.LayerNorm.bias", (D,)),
            (t + "embeddings.position_embeddings.weight", (cfg["max_position_embeddings"], D))]
    for i in range(cfg["num_hidden_layers"]):
        p = f"{t}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            spec += [(p + f"attention.self.{n}.weight", (D, D)), (p + f"attention.self.{n}.bias", (D,))]
        spec += [(p + "attention.output.dense.weight", (D, D)), (p + "attention.output.dense.bias", (D,)),
                 (p + "attention.output.LayerNorm.weight", (D,)), (p + "attention.output.LayerNorm.bias", (D,)),
                 (p + "intermediate.dense.weight", (I, D)), (p + "intermediate.dense.bias", (I,)),
                 (p + "output.dense.weight", (D, I)), (p + "output.dense.bias", (D,)),
                 (p + "output.LayerNorm.weight", (D,)), (p + "output.LayerNorm.bias", (D,))]
    spec += [(t + "pooler.dense.weight", (D, D)), (t + "pooler.dense.bias", (D,)),
             ("text_encoder.text_projection.linear1.weight", (P, D)), ("text_encoder.text_projection.linear1.bias", (P,)),
             ("text_encoder.text_projection.linear2.weight", (P, P)), ("text_encoder.text_projection.linear2.bias", (P,))]
    for i in range(num_projection_layers):
        p = f"text_projection.layers.{i}."
        din = P if i == 0 else emb_dim
        spec += [(p + "projection.weight", (emb_dim, din)), (p + "projection.bias", (emb_dim,)), (p + "fc.weight", (emb_dim, emb_dim)),
                 (p + "fc.bias", (emb_dim,)), (p + "layer_norm.weight", (emb_dim,)), (p + "layer_norm.bias", (emb_dim,))]
    return spec


def text_random_state_dict(cfg: Optional[dict] = None, num_projection_layers: int = 2, seed: int = 7) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights (no checkpoints offline): N(0, 0.02) matrices like the RoBERTa initialiser, perturbed norms."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in text_param_spec(cfg or CLAP_TEXT, num_projection_layers):
        if "LayerNorm" in name or "layer_norm" in name:
            sd[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if name.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            sd[name] = 0.02 * torch.randn(shape, generator=g)
        elif "embeddings" in name:
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            sd[name] = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
    return sd


class TextEncoder:
    """``get_text_features(input_ids, attention_mask) -> [B, emb_dim]`` fp32 on the device."""

    def __init__(self, cfg: Optional[dict] = None, num_projection_layers: int = 2, emb_dim: int = 512, device=None):
        self.cfg = dict(cfg or CLAP_TEXT)
        if self.cfg["hidden_size"] // self.cfg["num_attention_heads"] != 64:
            raise NotImplementedError("head size must be 64 (ClapTextConfig)")
        self.num_projection_layers, self.emb_dim = num_projection_layers, emb_dim
        self.device = torch.device(device if device is not None else "cuda")
        self._sd = None
        self._plans = {}

    def load_state_dict(self, state_dict, strict=True):
        spec = text_param_spec(self.cfg, self.num_projection_layers, self.emb_dim)
        missing = [k for k, _ in spec if k not in state_dict]
        if missing:
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:4]}...")
        for k, shp in spec:
            if tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {shp}")
        sd = OrderedDict((k, state_dict[k].detach().float().cpu().contiguous()) for k, _ in spec)
        self._sd, dev, t = sd, self.device, "text_encoder.text_model."
        f = lambda k: sd[k].to(dev)
        self.word, self.pos, self.type0 = f(t + "embeddings.word_embeddings.weight"), f(t + "embeddings.position_embeddings.weight"), \
            f(t + "embeddings.token_type_embeddings.weight")[0].contiguous()
        self.emb_g, self.emb_b = f(t + "embeddings.LayerNorm.weight"), f(t + "embeddings.LayerNorm.bias")
        self.layers = []
        for i in range(self.cfg["num_hidden_layers"]):
            p = f"{t}encoder.layer.{i}."
            wqkv = torch.cat([sd[p + f"attention.self.{n}.weight"] for n in ("query", "key", "value")])
            bqkv = torch.cat([sd[p + f"attention.self.{n}.bias"] for n in ("query", "key", "value")])
            conv = lambda w, b: pack_conv_s1(w[:, :, None, None], b).to(dev)
            self.layers.append(dict(qkv=conv(wqkv, bqkv), out=conv(sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]),
                                    ln1=(f(p + "attention.output.LayerNorm.weight"), f(p + "attention.output.LayerNorm.bias")),
                                    ffn1=conv(sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]),
                                    ffn2=conv(sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]),
                                    ln2=(f(p + "output.LayerNorm.weight"), f(p + "output.LayerNorm.bias"))))
        self.pool = (f(t + "pooler.dense.weight"), f(t + "pooler.dense.bias"))
        self.proj = [(f(f"text_encoder.text_projection.linear{j}.weight"), f(f"text_encoder.text_projection.linear{j}.bias")) for j in (1, 2)]
        self.head = []
        for i in range(self.num_projection_layers):
            p = f"text_projection.layers.{i}."
            self.head.append(dict(proj=(f(p + "projection.weight"), f(p + "projection.bias")), fc=(f(p + "fc.weight"), f(p + "fc.bias")),
                                  ln=(f(p + "layer_norm.weight"), f(p + "layer_norm.bias"))))
        self._plans.clear()
        return self

    def state_dict(self):
        return OrderedDict(self._sd)

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device:
            raise NotImplementedError("construct TextEncoder with device=... (weights are packed for one device)")
        return self

    # ---- one (B, L) plan: buffers + launch list ----
    def _plan(self, B: int, L: int):
        key = (B, L)
        if key in self._plans:
            return self._plans[key]
        lib, dev, cfg = _lib.load(), self.device, self.cfg
        D, I, Hh, T = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_attention_heads"], B * L
        stream = ops._stream
        Wt = 1
        while Wt < 128 and T % (2 * Wt) == 0:      # the token axis as an H x W grid of "pixels" for the 1x1-conv GEMM
            Wt *= 2
        Ht = T // Wt
        a16 = lambda c: torch.empty((1, Ht, Wt, c), dtype=ops.ACT, device=dev)
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        pl = dict(ids=torch.zeros((B, L), dtype=torch.long, device=dev), mask=torch.ones((B, L), dtype=torch.long, device=dev), ops=[], keep=[])
        h, h1, qkv, ctx, ff = a16(D), a16(D), a16(3 * D), a16(D), a16(I)
        run = pl["ops"]
        run.append(lambda: check(lib.ds_text_embed_ln(pl["ids"].data_ptr(), self.word.data_ptr(), self.pos.data_ptr(), self.type0.data_ptr(),
                                                      self.emb_g.data_ptr(), self.emb_b.data_ptr(), h.data_ptr(), B, L, D, cfg["pad_token_id"],
                                                      cfg["layer_norm_eps"], stream()), "text_embed_ln"))

        def gemm(pc, src, out, **kw):
            a, _, keep = conv_args(pc, src, None, 1, Ht, Wt, out=out, **kw)
            pl["keep"] += keep + [a]
            run.append(lambda a=a: run_conv(a))

        def ln(x, gb):
            run.append(lambda: check(lib.ds_layernorm_rows(x.data_ptr(), gb[0].data_ptr(), gb[1].data_ptr(), T, D, cfg["layer_norm_eps"], stream()), "layernorm_rows"))

        for Ly in self.layers:
            gemm(Ly["qkv"], h, qkv)
            run.append(lambda: check(lib.ds_text_attention(qkv.data_ptr(), pl["mask"].data_ptr(), ctx.data_ptr(), B, L, Hh, 64 ** -0.5, stream()), "text_attention"))
            gemm(Ly["out"], ctx, h1, residual=h)
            ln(h1, Ly["ln1"])
            gemm(Ly["ffn1"], h1, ff, act=1)
            gemm(Ly["ffn2"], ff, h, residual=h1)
            ln(h, Ly["ln2"])
        P, E = cfg["projection_dim"], self.emb_dim
        cls, pooled, p1, feat = f32(B, D), f32(B, D), f32(B, P), f32(B, P)
        run.append(lambda: check(lib.ds_cls_gather(h.data_ptr(), cls.data_ptr(), B, L, D, stream()), "cls_gather"))
        run.append(lambda: ops.linear(cls, self.pool[0], self.pool[1], pooled, act_out=3))
        run.append(lambda: ops.linear(pooled, self.proj[0][0], self.proj[0][1], p1, act_out=4))
        run.append(lambda: ops.linear(p1, self.proj[1][0], self.proj[1][1], feat))
        run.append(lambda: check(lib.ds_l2_normalize_rows(feat.data_ptr(), B, P, stream()), "l2_normalize"))
        x = feat
        for Hd in self.head:
            projected, y = f32(B, E), f32(B, E)
            run.append(lambda x=x, projected=projected, Hd=Hd: ops.linear(x, Hd["proj"][0], Hd["proj"][1], projected))
            run.append(lambda projected=projected, y=y, Hd=Hd: ops.linear(projected, Hd["fc"][0], Hd["fc"][1], y, act_in=1))
            run.append(lambda projected=projected, y=y, Hd=Hd: check(lib.ds_add_layernorm_rows_f32(
                y.data_ptr(), projected.data_ptr(), Hd["ln"][0].data_ptr(), Hd["ln"][1].data_ptr(), B, E, 1e-5, stream()), "add_layernorm"))
            x = y
        pl["out"], pl["hidden"] = x, h
        pl["keep"] += [h, h1, qkv, ctx, ff, cls, pooled, p1, feat]
        self._plans[key] = pl
        return pl

    @torch.no_grad()
    def get_text_features(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self._sd is None:
            raise RuntimeError("TextEncoder: load_state_dict() has not been called")
        B, L = input_ids.shape
        if L > self.cfg["max_position_embeddings"] - 2:
            raise RuntimeError(f"sequence length {L} exceeds the {self.cfg['max_position_embeddings'] - 2} positions of the text model")
        pl = self._plan(B, L)
        pl["ids"].copy_(input_ids.to(self.device, torch.long))
        if attention_mask is None:
            pl["mask"].fill_(1)
        else:
            pl["mask"].copy_(attention_mask.to(self.device, torch.long))
        for fn in pl["ops"]:
            fn()
        return pl["out"].clone()

    def num_launches(self, B: int, L: int) -> int:
        return len(self._plan(B, L)["ops"])

"""ConditionedUnet -- B200 drop-in for the reference's epsilon-predictor
(model/diffusion.py:21-258: ``ConditionedUnet(**unetConfig)``, ``.load_state_dict``, ``model(x, time, condition)``).

Same constructor arguments, same ``state_dict`` key names, same call signature and result
(fp32 NCHW epsilon).  The forward pass is a fixed sequence of calls into the C ABI
(ds_stem_im2col, ds_dwconv7, ds_conv_gemm, ds_attn_*, ds_gn_apply_residual, ds_linear ...);
activations live in HBM as bf16 NHWC and GroupNorm(1,C) never runs as its own pass -- its
statistics come out of the producing kernel's epilogue and its affine is folded into the
consuming convolution.  The call sequence for a given (N, H, W) is built once
("plan") and then replayed, which is what makes it capturable in a CUDA graph."""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib, engine, ops, weights as W
from ._lib import check
from .ops import PackedConv, Stats, conv_args, pack_conv_down, pack_conv_s1, pack_conv_up, run_conv

HEADS, DHEAD = W.ATTN_HEADS, W.ATTN_DIM_HEAD
HID = HEADS * DHEAD
# to_qkv + q soft-max + partial context as ONE tcgen05 kernel that keeps k and v on chip (ds_attn_qkv_ctx); False builds the
# plan from the separate ds_conv_gemm(to_qkv) + ds_attn_ctx_partial launches (kept for the A/B parity test).
FUSED_ATTN = True


class _Block:
    """Packed constants of one ConvNextBlock (diffusion_components.py:107-139)."""

    def __init__(self, sd, p, dim, dim_out, has_time):
        self.p, self.dim, self.dim_out = p, dim, dim_out
        self.dw = sd[p + "ds_conv.weight"].float().reshape(dim, 49).t().contiguous()      # [49][C]
        self.dw_bias = sd[p + "ds_conv.bias"].float()
        self.has_time = has_time
        self.conv1 = pack_conv_s1(sd[p + "net.1.weight"], sd[p + "net.1.bias"], sd[p + "net.0.weight"].float(), sd[p + "net.0.bias"].float())
        self.conv2 = pack_conv_s1(sd[p + "net.4.weight"], sd[p + "net.4.bias"], sd[p + "net.3.weight"].float(), sd[p + "net.3.bias"].float())
        self.res = pack_conv_s1(sd[p + "res_conv.weight"], sd[p + "res_conv.bias"]) if (p + "res_conv.weight") in sd else None
        self.t_off = 0   # column offset of this block's time projection in the fused [N, sum(dim)] buffer
        self.t_dim = dim

    def to(self, dev):
        self.dw = self.dw.to(dev)
        for c in (self.conv1, self.conv2, self.res):
            if c is not None:
                c.to(dev)
        return self


class _ResBlock:
    """Packed constants of one ResnetBlock (use_convnext=False; diffusion_components.py:59-104): two (conv3x3 -> GroupNorm(groups) ->
    SiLU) blocks with the time embedding added in between, plus the 1x1 residual projection (an identity matrix where the reference
    uses nn.Identity, so the residual add and the statistics of the block's output stay in the conv epilogue)."""

    def __init__(self, sd, p, dim, dim_out, has_time):
        self.p, self.dim, self.dim_out = p, dim, dim_out
        self.has_time = has_time
        self.t_dim = dim_out                          # width of this block's slice of the fused time projection
        self.dw_bias = torch.zeros(dim_out)            # (nothing to add to mlp.1.bias)
        self.conv1 = pack_conv_s1(sd[p + "block1.proj.weight"], sd[p + "block1.proj.bias"].float())
        self.conv2 = pack_conv_s1(sd[p + "block2.proj.weight"], sd[p + "block2.proj.bias"].float())
        self.g1, self.b1 = sd[p + "block1.norm.weight"].float().contiguous(), sd[p + "block1.norm.bias"].float().contiguous()
        self.g2, self.b2 = sd[p + "block2.norm.weight"].float().contiguous(), sd[p + "block2.norm.bias"].float().contiguous()
        if (p + "res_conv.weight") in sd:
            self.res = pack_conv_s1(sd[p + "res_conv.weight"], sd[p + "res_conv.bias"].float())
        else:
            self.res = pack_conv_s1(torch.eye(dim).reshape(dim, dim, 1, 1), torch.zeros(dim))
        self.t_off = 0

    def to(self, dev):
        for c in (self.conv1, self.conv2, self.res):
            c.to(dev)
        self.g1, self.b1, self.g2, self.b2 = self.g1.to(dev), self.b1.to(dev), self.g2.to(dev), self.b2.to(dev)
        return self


class _Attn:
    """Packed constants of Residual(PreNorm(LinearCrossAttentionAdd)) (diffusion_components.py:252-293,142-152)."""

    def __init__(self, sd, p, dim):
        self.p, self.dim = p, dim
        self.qkv = pack_conv_s1(sd[p + "fn.fn.to_qkv.weight"], None, sd[p + "fn.norm.weight"].float(), sd[p + "fn.norm.bias"].float())
        self.wout = sd[p + "fn.fn.to_out.0.weight"].float().reshape(dim, HID).contiguous()
        cout_pad = ops.pad16(dim)
        # the per-sample matrix M = Wout . ctx is produced on the device; this PackedConv only carries e2 (= bias), taps, sizes
        e2 = torch.zeros(1, cout_pad)
        e2[0, :dim] = sd[p + "fn.fn.to_out.0.bias"].float()
        self.out = PackedConv(weight=torch.zeros(1, dtype=ops.ACT), e2=e2, e1=None, taps=[[(0, 0, 0)]], cin=HID, cout=dim,
                              cout_pad=cout_pad, ncls=1, kind="s1")
        self.gamma = sd[p + "fn.fn.to_out.1.weight"].float().contiguous()
        self.beta = sd[p + "fn.fn.to_out.1.bias"].float().contiguous()
        self.c_off = 0   # column offset of this site's [label_query | label_key | 0] bias in the fused cond buffer

    def to(self, dev):
        self.qkv.to(dev)
        self.out.to(dev)
        self.wout, self.gamma, self.beta = self.wout.to(dev), self.gamma.to(dev), self.beta.to(dev)
        return self


class ConditionedUnet:
    def __init__(self, in_dim, out_dim=None, down_dims=None, up_dims=None, mid_depth=3, with_time_emb=True, time_dim=None,
                 resnet_block_groups=8, use_convnext=True, convnext_mult=2, attn_type="linear_cat", n_label_class=11,
                 condition_type="instrument_family", label_emb_dim=128, device=None, batch_invariant=False):
        if attn_type not in ("linear_cat", "linear_add"):
            raise NotImplementedError()                       # diffusion.py:96
        if condition_type not in ("instrument_family", "natural_language_prompt"):
            raise NotImplementedError()                       # diffusion_components.py:165
        if up_dims is None:
            up_dims = [128, 128, 64, 32]
        if down_dims is None:
            down_dims = [32, 32, 64, 128]
        assert len(down_dims) == len(up_dims), "len(down_dims) != len(up_dims)"
        assert down_dims[0] == up_dims[-1], "down_dims[0] != up_dims[-1]"
        assert up_dims[0] == down_dims[-1], "up_dims[0] != down_dims[-1]"
        self.cfg = W.unet_config(in_dim=in_dim, out_dim=out_dim, down_dims=list(down_dims), up_dims=list(up_dims), mid_depth=mid_depth,
                                 time_dim=time_dim, convnext_mult=convnext_mult, attn_type=attn_type, condition_type=condition_type,
                                 label_emb_dim=label_emb_dim, use_convnext=bool(use_convnext), resnet_block_groups=int(resnet_block_groups),
                                 with_time_emb=bool(with_time_emb), n_label_class=int(n_label_class))
        # (not a reference argument) False: jobs too small to fill the SMs run narrow N tiles -- lower latency, but a sample's low-order bits
        # then depend on the batch it runs in (grouping of the GroupNorm partial sums); True: always the widest tiling, bit-identical
        # samples whatever the batch / shard (ds_unet_config.batch_invariant)
        self.cfg["batch_invariant"] = bool(batch_invariant)
        for d in set(self.cfg["down_dims"] + self.cfg["up_dims"]):
            if d % 32:
                raise NotImplementedError(f"channel widths must be multiples of 32 (got {d})")
        self.device = torch.device(device if device is not None else "cuda")
        self._sd: Optional[OrderedDict] = None
        self._plans: Dict[Tuple, "_Plan"] = {}
        self.weights_version = 0      # bumped by every repack; graph caches key on it (sampler._run_graph_loop)
        self._engine = None
        self.use_engine = True        # False: route forward() / the sampling graph through the operator-level Python plan (A/B tests)
        self.training = False

    # ---- nn.Module-like surface ------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True):
        spec = W.unet_param_spec(self.cfg)
        missing = [k for k, _ in spec if k not in state_dict]
        unexpected = [k for k in state_dict if k not in dict(spec)]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:4]}..., unexpected {unexpected[:4]}...")
        for k, shp in spec:
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {shp}")
        self._sd = OrderedDict((k, state_dict[k].detach().float().cpu().contiguous()) for k, _ in spec)
        self._pack()
        self._plans.clear()
        return self

    def state_dict(self):
        return OrderedDict(self._sd)

    def parameters(self):
        return iter([self._probe])

    def eval(self):
        return self

    def to(self, device):
        device = torch.device(device)
        if device != self.device:
            self.device = device
            if self._sd is not None:
                self._pack()
                self._plans.clear()
        return self

    def __call__(self, x, time, condition=None):
        return self.forward(x, time, condition)

    # ---- packing -----------------------------------------------------------------------------
    def _pack(self):
        sd, cfg, dev = self._sd, self.cfg, self.device
        dd, ud, td, L = cfg["down_dims"], cfg["up_dims"], cfg["time_dim"], cfg["label_emb_dim"]
        self.weights_version += 1
        self._probe = torch.zeros(1, device=dev)
        # the module-level C ABI handle (ds_unet_*): forward() and the sampling graph run through it for the deployed family;
        # the operator-level plan below stays for the other variants, the per-layer taps of the parity tests and bench.py's
        # per-kernel timings
        self._engine = engine.UnetEngine(cfg, sd, dev) if engine.unet_supported(cfg) else None
        self.blocks: "OrderedDict[str, _Block]" = OrderedDict()
        self.attns: "OrderedDict[str, _Attn]" = OrderedDict()
        self.samplers: Dict[str, PackedConv] = {}

        def blk(p, dim, dim_out, has_time=True):
            self.blocks[p] = (_Block if cfg["use_convnext"] else _ResBlock)(sd, p, dim, dim_out, has_time and cfg["with_time_emb"])

        def att(p, dim):
            self.attns[p] = _Attn(sd, p, dim)

        skips = []
        for i, (cin, cout) in enumerate(zip(dd[:-1], dd[1:])):
            p = f"downs.{i}."
            blk(p + "0.", cin, cout); att(p + "1.", cout); blk(p + "2.", cout, cout); att(p + "3.", cout)
            self.samplers[p + "4."] = pack_conv_down(sd[p + "4.weight"], sd[p + "4.bias"].float())
            skips.append(cout)
        mid = dd[-1]
        for j in range(cfg["mid_depth"] - 1):
            blk(f"mid_left.{j}.", mid, mid)
        blk("mid_mid.0.", mid, mid); att("mid_mid.1.", mid); blk("mid_mid.2.", mid, mid)
        for j in range(cfg["mid_depth"] - 1):
            blk(f"mid_right.{j}.", 2 * mid, mid)
        for i, (cin, cout) in enumerate(zip(ud[:-1], ud[1:])):
            s = skips.pop()
            p = f"ups.{i}."
            blk(p + "0.", cin + s, cin); att(p + "1.", cin)
            self.samplers[p + "2."] = pack_conv_up(sd[p + "2.weight"], sd[p + "2.bias"].float())
            blk(p + "3.", cin + s, cout); att(p + "4.", cout)
            blk(p + "5.", cout + s, cout); att(p + "6.", cout)
        blk("final_conv.0.", dd[0] + ud[-1], ud[-1], has_time=False)
        self.final = pack_conv_s1(sd["final_conv.1.weight"], sd["final_conv.1.bias"].float())

        # fused time projection: rows = concat over blocks of mlp.1 (GELU on the input); bias += ds_conv.bias.
        rows, biases, off = [], [], 0
        for b in self.blocks.values():
            b.t_off = off
            if b.has_time:
                rows.append(sd[b.p + "mlp.1.weight"].float())
                biases.append(sd[b.p + "mlp.1.bias"].float() + b.dw_bias)
            else:
                rows.append(torch.zeros(b.t_dim, td))
                biases.append(b.dw_bias.clone())
            off += b.t_dim
        self.t_total = off
        self.t_w, self.t_b = torch.cat(rows).contiguous().to(dev), torch.cat(biases).contiguous().to(dev)
        # fused condition projection: per attention site [label_query | label_key | zeros(v)] (added to q, k in the to_qkv epilogue),
        # or for "linear_cat" [label_key | label_value | zeros] (the extra key / value token, consumed by ds_attn_finalize_cat)
        rows, biases, off = [], [], 0
        first, second = ("label_query", "label_key") if cfg["attn_type"] == "linear_add" else ("label_key", "label_value")
        for a in self.attns.values():
            a.c_off = off
            rows += [sd[a.p + f"fn.fn.{first}.weight"].float(), sd[a.p + f"fn.fn.{second}.weight"].float(), torch.zeros(HID, L)]
            biases += [sd[a.p + f"fn.fn.{first}.bias"].float(), sd[a.p + f"fn.fn.{second}.bias"].float(), torch.zeros(HID)]
            off += 3 * HID
        self.c_total = off
        self.c_w, self.c_b = torch.cat(rows).contiguous().to(dev), torch.cat(biases).contiguous().to(dev)
        self.lab_w = sd["label_embedding.embedding.weight"].float().contiguous().to(dev)      # Linear weight, or the nn.Embedding table
        self.lab_b = sd["label_embedding.embedding.bias"].float().contiguous().to(dev) if "label_embedding.embedding.bias" in sd else None
        if cfg["with_time_emb"]:
            self.tm1_w, self.tm1_b = sd["time_mlp.1.weight"].float().contiguous().to(dev), sd["time_mlp.1.bias"].float().contiguous().to(dev)
            self.tm3_w, self.tm3_b = sd["time_mlp.3.weight"].float().contiguous().to(dev), sd["time_mlp.3.bias"].float().contiguous().to(dev)
        # stem as a GEMM over im2col patches: [Cout, Cin, 7, 7] -> [Cout, ky*32 + kx*4 + ci] (8th pixel slot and ci >= Cin are zero)
        w0 = sd["init_conv.weight"].float()
        wst = torch.zeros(dd[0], 7, 8, 4)
        wst[:, :, :7, :cfg["in_dim"]] = w0.permute(0, 2, 3, 1)
        cout_pad = ops.pad16(dd[0])
        e2 = torch.zeros(1, cout_pad)
        e2[0, :dd[0]] = sd["init_conv.bias"].float()
        wp = torch.zeros(1, cout_pad, 224, dtype=ops.ACT)
        wp[0, :dd[0]] = wst.reshape(dd[0], 224).to(ops.ACT)
        self.stem = PackedConv(weight=wp.contiguous(), e2=e2, e1=None, taps=[[(0, 0, 0)]], cin=224, cout=dd[0], cout_pad=cout_pad,
                               ncls=1, kind="s1").to(dev)
        for m in list(self.blocks.values()) + list(self.attns.values()):
            m.to(dev)
        for s in self.samplers.values():
            s.to(dev)
        self.final.to(dev)

    # ---- forward -----------------------------------------------------------------------------
    def plan(self, N: int, H: int, Wd: int, x_batch_mod: int = 0, uniform_time: bool = False) -> "_Plan":
        key = (N, H, Wd, x_batch_mod, uniform_time)
        if key not in self._plans:
            if self._sd is None:
                raise RuntimeError("ConditionedUnet: load_state_dict() has not been called")
            self._plans[key] = _Plan(self, N, H, Wd, x_batch_mod, uniform_time)
        return self._plans[key]

    @torch.no_grad()
    def forward(self, x, time, condition=None, taps: Optional[dict] = None):
        """x [N,4,H,W] fp32, time [N] int64, condition [N, label_emb_dim] fp32 -> eps [N,4,H,W] fp32 (diffusion.py:187-258)."""
        N, Cin, H, Wd = x.shape
        assert Cin == self.cfg["in_dim"]
        n_stage = len(self.cfg["down_dims"]) - 1
        if (H >> n_stage) < 1 or (Wd >> n_stage) < 1:
            raise RuntimeError(f"H={H}, W={Wd}: the map vanishes after {n_stage} stride-2 stages")      # torch: "Output size is too small"
        if (H >> n_stage) < 2 or (Wd >> n_stage) < 2:
            # the GroupNorm fold's border classes (top / middle / bottom x left / middle / right) assume every level is >= 2 pixels
            # in both directions: a 1-pixel level is at the same time a first and a last row / column
            raise NotImplementedError(f"H={H}, W={Wd}: a level of the U-Net is 1 pixel wide or high (minimum input {2 << n_stage} x {2 << n_stage})")
        if self._engine is not None and taps is None and self.use_engine:
            cdt = torch.long if self.cfg["condition_type"] == "instrument_family" else torch.float32
            return self._engine.forward(x.to(self.device, torch.float32).contiguous(), time.to(self.device, torch.long).contiguous(),
                                        None if condition is None else condition.to(self.device, cdt).contiguous())
        if condition is None and self.cfg["attn_type"] != "linear_add":
            raise NotImplementedError("condition=None with attn_type='linear_cat' runs through the module-level entry points only "
                                      "(the operator-level plan used for per-layer taps has the extra key / value token built in)")
        pl = self.plan(N, H, Wd)
        pl.x.copy_(x.to(self.device, torch.float32))
        pl.t.copy_(time.to(self.device, torch.long))
        if condition is None:
            pl.sbias.zero_()          # diffusion.py:199-202, diffusion_components.py:279-283: no label_query / label_key terms at all
        else:
            pl.cond.copy_(condition.to(self.device, pl.cond.dtype).reshape(pl.cond.shape))
            pl.run_cond()
        pl.run()
        if taps is not None:
            pl.export_taps(taps)
        return pl.eps.clone()


class _Plan:
    """The static call sequence + buffers of one (N, H, W) U-Net evaluation."""

    def __init__(self, net: ConditionedUnet, N: int, H: int, Wd: int, x_batch_mod: int, uniform_time: bool):
        self.net, self.N, self.H, self.W = net, N, H, Wd
        dev, cfg = net.device, net.cfg
        lib = _lib.load()
        f32 = dict(dtype=torch.float32, device=dev)
        self.keep: list = []
        self.ops: List[Tuple[str, Callable[[], None]]] = []
        self.cond_ops: List[Tuple[str, Callable[[], None]]] = []
        self.named: Dict[str, Tuple[torch.Tensor, int]] = {}
        nb = x_batch_mod if x_batch_mod > 0 else N
        self.x = torch.zeros((nb, cfg["in_dim"], H, Wd), **f32)
        self.t = torch.zeros((N,), dtype=torch.long, device=dev)
        labels = cfg["condition_type"] == "instrument_family"       # integer class labels through nn.Embedding (diffusion_components.py:161)
        self.cond = torch.zeros((N,), dtype=torch.long, device=dev) if labels else torch.zeros((N, cfg["label_emb_dim"]), **f32)
        self.eps = torch.zeros((N, cfg["out_dim"], H, Wd), **f32)
        dd, td = cfg["down_dims"], cfg["time_dim"]
        self.uniform_time = uniform_time
        NT = 1 if uniform_time else N          # all samples share t inside the sampling loop
        self.t_stride = 0 if uniform_time else net.t_total
        stream = ops._stream

        def act(n, h, w, c):
            return torch.empty((n, h, w, c), dtype=ops.ACT, device=dev)

        scratch: Dict[Tuple, torch.Tensor] = {}

        def scr(role, n, h, w, c):
            k = (role, n, h, w, c)
            if k not in scratch:
                scratch[k] = act(n, h, w, c)
            return scratch[k]

        self.meta: Dict[str, dict] = {}       # per op: kernel family + algorithmic FLOPs / HBM bytes (bench.py's per-kernel rooflines)

        def add(name, fn, family="misc", flops=0.0, bytes=0.0):
            self.ops.append((name, fn))
            self.meta[name] = dict(family=family, flops=float(flops), bytes=float(bytes))

        # ---- condition path (step-invariant: run once per sample() call) ----
        cemb = torch.empty((N, cfg["label_emb_dim"]), **f32)
        self.sbias = torch.empty((N, net.c_total), **f32)
        if labels:
            self.cond_ops.append(("label_embedding", lambda: check(lib.ds_embedding_gather(net.lab_w.data_ptr(), self.cond.data_ptr(), cemb.data_ptr(), N,
                                                                                           cfg["label_emb_dim"], net.lab_w.shape[0], stream()), "embedding_gather")))
        else:
            self.cond_ops.append(("label_embedding", lambda: ops.linear(self.cond, net.lab_w, net.lab_b, cemb)))
        self.cond_ops.append(("label_qk", lambda: ops.linear(cemb, net.c_w, net.c_b, self.sbias)))
        # ---- time path ----
        sin = torch.empty((NT, dd[0]), **f32)
        t1 = torch.empty((NT, td), **f32)
        temb = torch.empty((NT, td), **f32)
        self.tbias = torch.empty((NT, net.t_total), **f32)
        if cfg["with_time_emb"]:
            add("time_sin", lambda: check(lib.ds_sinusoidal_embedding(self.t.data_ptr(), sin.data_ptr(), NT, dd[0], stream()), "sinusoidal"))
            add("time_mlp1", lambda: ops.linear(sin, net.tm1_w, net.tm1_b, t1, act_out=1))
            add("time_mlp3", lambda: ops.linear(t1, net.tm3_w, net.tm3_b, temb))
            add("time_proj", lambda: ops.linear(temb, net.t_w, net.t_b, self.tbias, act_in=1 if cfg["use_convnext"] else 2))
            self.named["time_emb"] = (temb, -1)
        else:
            # with_time_emb=False (diffusion.py:107-109,211): no time terms; the per-channel bias of each block's first conv is all
            # the fused buffer carries (one row, read with stride 0)
            self.tbias = net.t_b.reshape(1, -1).clone()
            self.t_stride = 0

        def conv(name, pc, s0, s1, h, w, n=None, **kw):
            a, st, keep = conv_args(pc, s0, s1, N if n is None else n, h, w, wide_tiles=cfg["batch_invariant"], **kw)
            self.keep += keep + [a]
            fam, fl, by = ops.conv_cost(a, pc)
            add(name, lambda a=a: run_conv(a), fam, fl, by)
            return st

        def block(p, s0, s1, h, w, n=N, s0_mod=0):
            """One ConvNextBlock over n samples; s0_mod > 0: source 0 holds s0_mod samples shared by the guidance halves."""
            b = net.blocks[p]
            hbuf = scr("dw", n, h, w, b.dim)
            tb = self.tbias[:, b.t_off:]
            st_h = ops.dwconv7_stats(n, b.dim, h, w, dev)
            add(p + "ds_conv", lambda: ops.dwconv7(s0, s1, b.dw, tb, self.t_stride, hbuf, n, h, w, stats=st_h, src_batch_mod=s0_mod),
                "dwconv7", 2.0 * 49 * n * h * w * b.dim, 2.0 * n * h * w * b.dim * 2)
            y = scr("hid", n, h, w, b.conv1.cout)
            st_y = conv(p + "net.1", b.conv1, hbuf, None, h, w, n=n, out=y, stats_in=st_h, act=1, want_stats=True)
            if b.res is not None:
                r = scr("res", n, h, w, b.dim_out)
                conv(p + "res_conv", b.res, s0, s1, h, w, n=n, out=r, src_batch_mod=s0_mod)
            else:
                assert s1 is None and s0_mod == 0
                r = s0
            o = act(n, h, w, b.dim_out)
            st_o = conv(p + "net.4", b.conv2, y, None, h, w, n=n, out=o, stats_in=st_y, residual=r, want_stats=True)
            self.named[p[:-1]] = (o, b.dim_out)
            return o, st_o

        GN_CHUNKS = 32

        def gn_silu(name, x, c, h, w, n, gamma, beta, role):
            """GroupNorm(groups, eps 1e-5) + SiLU as statistics + apply passes (the VQGAN GroupNorm kernels)."""
            G = cfg["resnet_block_groups"]
            part = torch.empty((n, G, GN_CHUNKS, 2), **f32)
            out = scr(role, n, h, w, c)
            add(name + ".stats", lambda: check(lib.ds_group_stats(x.data_ptr(), part.data_ptr(), n, c, c, G, h * w, GN_CHUNKS, stream()), "group_stats"))
            add(name + ".apply", lambda: check(lib.ds_gn_act(x.data_ptr(), out.data_ptr(), part.data_ptr(), GN_CHUNKS, gamma.data_ptr(),
                                                            beta.data_ptr(), n, c, c, G, h * w, 1e-5, 2, stream()), "gn_act"))
            self.keep.append(part)
            return out

        def res_block(p, s0, s1, h, w, n=N, s0_mod=0):
            """One ResnetBlock over n samples (diffusion_components.py:79-104)."""
            b = net.blocks[p]
            y1 = scr("rb_y1", n, h, w, b.dim_out)
            conv(p + "block1.proj", b.conv1, s0, s1, h, w, n=n, out=y1, src_batch_mod=s0_mod)
            h1 = gn_silu(p + "block1.norm", y1, b.dim_out, h, w, n, b.g1, b.b1, "rb_h1")
            if b.has_time:
                tb = self.tbias[:, b.t_off:]
                add(p + "temb", lambda: check(lib.ds_add_channel_bias(h1.data_ptr(), tb.data_ptr(), self.t_stride, n, b.dim_out, h * w, stream()),
                                              "add_channel_bias"))
            y2 = scr("rb_y2", n, h, w, b.dim_out)
            conv(p + "block2.proj", b.conv2, h1, None, h, w, n=n, out=y2)
            h2 = gn_silu(p + "block2.norm", y2, b.dim_out, h, w, n, b.g2, b.b2, "rb_h2")
            o = act(n, h, w, b.dim_out)
            st_o = conv(p + "res_conv", b.res, s0, s1, h, w, n=n, out=o, residual=h2, want_stats=True, src_batch_mod=s0_mod)
            self.named[p[:-1]] = (o, b.dim_out)
            return o, st_o

        if not cfg["use_convnext"]:
            block = res_block

        def attn(p, x, st_x, h, w, x_mod=0):
            """x_mod > 0: x (and its statistics) hold x_mod samples shared by the guidance halves."""
            a = net.attns[p]
            npix = h * w
            sb = self.sbias[:, a.c_off:a.c_off + 3 * HID]
            cat = cfg["attn_type"] == "linear_cat"
            qp = scr("qp", N, h, w, HID)
            part = torch.empty((lib.ds_attn_part_floats(N, HEADS, npix),), **f32)
            M = torch.empty((N, a.out.cout_pad, HID), dtype=ops.ACT, device=dev)
            if FUSED_ATTN:
                n_in = x_mod if x_mod > 0 else N
                add(p + "qkv_ctx", lambda: check(lib.ds_attn_qkv_ctx(
                    x.data_ptr(), a.dim, x_mod, st_x.buf.data_ptr(), st_x.slots, a.qkv.weight.data_ptr(), a.qkv.e1.data_ptr(), a.qkv.e2.data_ptr(),
                    None if cat else sb.data_ptr(), self.sbias.stride(0), qp.data_ptr(), part.data_ptr(), N, HEADS, npix, float(DHEAD ** -0.5),
                    stream()), "attn_qkv_ctx"),
                    "attn_qkv_ctx", 2.0 * N * npix * a.dim * 3 * HID + 2.0 * N * HEADS * DHEAD * DHEAD * npix,
                    (n_in * npix * a.dim + N * npix * HID + 3 * HID * a.dim) * 2.0)
            else:
                qkv = scr("qkv", N, h, w, 3 * HID)
                conv(p + "to_qkv", a.qkv, x, None, h, w, out=qkv, stats_in=st_x, sbias=None if cat else sb, src_batch_mod=x_mod)
                add(p + "ctx", lambda: check(lib.ds_attn_ctx_partial(qkv.data_ptr(), qp.data_ptr(), part.data_ptr(), N, HEADS, npix, 0,
                                                                      float(DHEAD ** -0.5), stream()), "attn_ctx_partial"),
                    "attn_ctx", 2.0 * N * HEADS * DHEAD * DHEAD * npix, N * npix * (3 * HID + HID) * 2.0)
            if cat:
                lk, lv = sb[:, :HID], sb[:, HID:2 * HID]
                add(p + "fin", lambda: check(lib.ds_attn_finalize_cat(part.data_ptr(), lk.data_ptr(), lv.data_ptr(), self.sbias.stride(0),
                                                                       a.wout.data_ptr(), M.data_ptr(), N, HEADS, npix, a.dim, a.out.cout_pad,
                                                                       stream()), "attn_finalize_cat"), "attn_finalize", 0.0, part.numel() * 4.0)
            else:
                add(p + "fin", lambda: check(lib.ds_attn_finalize(part.data_ptr(), a.wout.data_ptr(), M.data_ptr(), N, HEADS, npix, a.dim,
                                                                   a.out.cout_pad, stream()), "attn_finalize"), "attn_finalize", 0.0, part.numel() * 4.0)
            y = scr("atty", N, h, w, a.dim)
            st_y = conv(p + "to_out", a.out, qp, None, h, w, out=y, want_stats=True, weight_override=M, per_sample_weights=True)
            o = act(N, h, w, a.dim)
            add(p + "gn_res", lambda: check(lib.ds_gn_apply_residual(y.data_ptr(), x.data_ptr(), o.data_ptr(), st_y.buf.data_ptr(), st_y.slots,
                                                                      a.gamma.data_ptr(), a.beta.data_ptr(), N, a.dim, npix, x_mod, stream()),
                                            "gn_apply_residual"), "gn_apply_residual", 0.0, 3.0 * N * npix * a.dim * 2)
            self.keep += [part, M]
            self.named[p[:-1]] = (o, a.dim)
            return o

        # ---- network (diffusion.py:187-258) ----
        n_stage = len(dd) - 1
        h, w = H, Wd
        # Classifier-free guidance inside the sampling loop: both halves of the doubled batch share the latent and the timestep
        # and differ only through the condition, which first enters in downs.0.1 (label_query / label_key).  init_conv and
        # downs.0.0 are therefore evaluated once for the nb distinct latents and read with a batch modulus afterwards.
        shared = nb if (x_batch_mod > 0 and uniform_time and nb < N) else 0
        self.shared = shared
        n0 = shared if shared else N
        x0 = act(n0, h, w, dd[0])
        col = act(nb, h, w, 224)
        add("init_im2col", lambda: check(lib.ds_stem_im2col(self.x.data_ptr(), col.data_ptr(), nb, cfg["in_dim"], H, Wd, stream()), "stem_im2col"),
            "stem_im2col", 0.0, nb * H * Wd * (cfg["in_dim"] * 4.0 + 224 * 2.0))
        conv("init_conv", net.stem, col, None, h, w, n=n0, out=x0, src_batch_mod=0 if shared else x_batch_mod)
        self.named["init_conv"] = (x0, dd[0])
        hs = [x0]
        sizes: List[Tuple[int, int]] = []
        x = x0
        for i in range(n_stage):
            p = f"downs.{i}."
            if i == 0 and shared:
                x, st = block(p + "0.", x, None, h, w, n=shared)
                x = attn(p + "1.", x, st, h, w, x_mod=shared); hs.append(x)
            else:
                x, st = block(p + "0.", x, None, h, w)
                x = attn(p + "1.", x, st, h, w); hs.append(x)
            x, st = block(p + "2.", x, None, h, w)
            x = attn(p + "3.", x, st, h, w); hs.append(x)
            pc = net.samplers[p + "4."]
            d = act(N, h // 2, w // 2, pc.cout)
            conv(p + "4", pc, x, None, h, w, out=d)
            self.named[p + "4"] = (d, pc.cout)
            sizes.append((h, w))
            x = d; h //= 2; w //= 2; hs.append(x)
        for j in range(cfg["mid_depth"] - 1):
            x, _ = block(f"mid_left.{j}.", x, None, h, w); hs.append(x)
        x, st = block("mid_mid.0.", x, None, h, w)
        x = attn("mid_mid.1.", x, st, h, w)
        x, _ = block("mid_mid.2.", x, None, h, w)
        for j in range(cfg["mid_depth"] - 1):
            x, _ = block(f"mid_right.{j}.", hs.pop(), x, h, w)
        for i in range(n_stage):
            p = f"ups.{i}."
            x, st = block(p + "0.", hs.pop(), x, h, w)
            x = attn(p + "1.", x, st, h, w)
            pc = net.samplers[p + "2."]
            hp, wp = sizes.pop()                  # the skip's size: 2h / 2w, or one more where the level was odd (pad_to_match)
            if (hp, wp) == (2 * h, 2 * w):
                u = act(N, hp, wp, pc.cout)
            else:
                u = torch.zeros((N, hp, wp, pc.cout), dtype=ops.ACT, device=dev)      # the conv never writes the padding
            conv(p + "2", pc, x, None, h, w, out=u)
            self.named[p + "2"] = (u, pc.cout)
            x = u; h, w = hp, wp
            x, st = block(p + "3.", hs.pop(), x, h, w)
            x = attn(p + "4.", x, st, h, w)
            x, st = block(p + "5.", hs.pop(), x, h, w)
            x = attn(p + "6.", x, st, h, w)
        x, _ = block("final_conv.0.", hs.pop(), x, h, w, s0_mod=shared)       # the last skip is init_conv's output
        conv("final_conv.1", net.final, x, None, h, w, out_f32=self.eps)
        self.keep.append(scratch)

    def run_cond(self):
        for _, fn in self.cond_ops:
            fn()

    def run(self):
        for _, fn in self.ops:
            fn()

    def num_launches(self) -> int:
        """Kernels launched by one run(): one per op, two for the attention finalize (reduce + fold)."""
        return len(self.ops) + sum(1 for name, _ in self.ops if name.endswith("fin"))

    def export_taps(self, taps: dict):
        """Named intermediates as fp32 NCHW (for per-layer parity tests)."""
        for name, (t, c) in self.named.items():
            if c < 0:
                taps[name] = t.clone()
            else:
                v = t.float().permute(0, 3, 1, 2).contiguous()
                if v.shape[0] < self.N:                      # evaluated once for both guidance halves
                    v = v.repeat(self.N // v.shape[0], 1, 1, 1)
                taps[name] = v

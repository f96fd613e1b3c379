"""ctypes hosts of the module-level C ABI (include/diffusynth_b200.h, "Module-level entry points"): the U-Net, the VQGAN and the
sampling graph as opaque handles inside the library (csrc/engine.cu).  The reference-facing classes (ConditionedUnet, VQGAN,
DiffSynthSampler) hold one of these and forward their calls; nothing here computes."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import DsError, check

DS_MAX_LEVELS = 8


class UnetConfig(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("out_dim", C.c_int32), ("n_levels", C.c_int32),
                ("down_dims", C.c_int32 * DS_MAX_LEVELS), ("up_dims", C.c_int32 * DS_MAX_LEVELS),
                ("mid_depth", C.c_int32), ("with_time_emb", C.c_int32), ("time_dim", C.c_int32), ("use_convnext", C.c_int32),
                ("convnext_mult", C.c_int32), ("attn_type", C.c_int32), ("condition_type", C.c_int32), ("label_emb_dim", C.c_int32),
                ("n_label_class", C.c_int32), ("resnet_block_groups", C.c_int32), ("batch_invariant", C.c_int32)]


class UnetPlanIO(C.Structure):
    _fields_ = [("plan", C.c_int32), ("d_cond", C.c_void_p), ("d_eps", C.c_void_p), ("launches", C.c_int32), ("cond_launches", C.c_int32)]


class VqganConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("embedding_dim", C.c_int32),
                ("n_hidden", C.c_int32), ("hidden_channels", C.c_int32 * DS_MAX_LEVELS), ("block_depth", C.c_int32),
                ("n_attn_pos", C.c_int32), ("attn_pos", C.c_int32 * DS_MAX_LEVELS), ("attn_with_skip", C.c_int32),
                ("act_relu", C.c_int32), ("num_embeddings", C.c_int32), ("num_groups", C.c_int32), ("batch_invariant", C.c_int32)]


class SampleBuffers(C.Structure):
    _fields_ = [("d_imgs", C.c_void_p), ("d_coef", C.c_void_p), ("d_ttab", C.c_void_p), ("d_noise", C.c_void_p), ("d_cond", C.c_void_p),
                ("d_guide", C.c_void_p), ("d_init_noise", C.c_void_p), ("d_masks", C.c_void_p), ("d_blend_coef", C.c_void_p),
                ("d_quantized", C.c_void_p), ("d_indices", C.c_void_p), ("d_spec", C.c_void_p), ("d_wave", C.c_void_p)]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _load_params(fn, handle, state_dict) -> None:
    for name, t in state_dict.items():
        t = t.detach().to(torch.float32).contiguous().cpu()
        shape = (C.c_longlong * max(1, t.dim()))(*t.shape)
        check(fn(handle, name.encode(), t.data_ptr(), shape, t.dim()), f"load({name})")


def unet_supported(cfg: dict) -> bool:
    """Every constructor variant of the reference runs through the module-level entry points (limits: in_dim <= 4, <= 8 levels)."""
    return cfg["in_dim"] <= 4 and len(cfg["down_dims"]) <= DS_MAX_LEVELS


class UnetEngine:
    """ds_unet handle: ConditionedUnet behind the C ABI."""

    def __init__(self, cfg: dict, state_dict, device: torch.device):
        self.lib = _lib.load()
        self.cfg, self.device = cfg, device
        c = UnetConfig()
        c.in_dim, c.out_dim, c.n_levels = cfg["in_dim"], cfg["out_dim"], len(cfg["down_dims"])
        for i, (d, u) in enumerate(zip(cfg["down_dims"], cfg["up_dims"])):
            c.down_dims[i], c.up_dims[i] = d, u
        c.mid_depth, c.with_time_emb, c.time_dim = cfg["mid_depth"], int(cfg.get("with_time_emb", True)), cfg["time_dim"]
        c.use_convnext, c.resnet_block_groups = int(cfg.get("use_convnext", True)), int(cfg.get("resnet_block_groups", 8))
        c.convnext_mult, c.attn_type, c.label_emb_dim = cfg["convnext_mult"], int(cfg["attn_type"] == "linear_cat"), cfg["label_emb_dim"]
        c.condition_type, c.n_label_class = int(cfg["condition_type"] == "instrument_family"), cfg.get("n_label_class", 11)
        c.batch_invariant = int(bool(cfg.get("batch_invariant", False)))
        self.h = C.c_void_p()
        with torch.cuda.device(device):
            check(self.lib.ds_unet_create(C.byref(c), C.byref(self.h)), "ds_unet_create")
            _load_params(self.lib.ds_unet_load, self.h, state_dict)
            check(self.lib.ds_unet_finalize(self.h), "ds_unet_finalize")

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.ds_unet_destroy(h)
            except Exception:
                pass

    def forward(self, x: torch.Tensor, t: torch.Tensor, cond: Optional[torch.Tensor]) -> torch.Tensor:
        N, _, H, W = x.shape
        out = torch.empty((N, self.cfg["out_dim"], H, W), dtype=torch.float32, device=self.device)
        check(self.lib.ds_unet_forward(self.h, x.data_ptr(), t.data_ptr(), _ptr(cond), out.data_ptr(), N, H, W, _stream()), "ds_unet_forward")
        return out


class VqganEngine:
    """ds_vqgan handle: quantiser + Decoder + Encoder behind the C ABI."""

    def __init__(self, cfg: dict, state_dict, device: torch.device):
        self.lib = _lib.load()
        self.cfg, self.device = cfg, device
        c = VqganConfig()
        c.in_channels, c.out_channels, c.embedding_dim = cfg["in_channels"], cfg["out_channels"], cfg["embedding_dim"]
        c.n_hidden = len(cfg["hidden_channels"])
        for i, v in enumerate(cfg["hidden_channels"]):
            c.hidden_channels[i] = v
        c.block_depth, c.n_attn_pos = cfg["block_depth"], len(cfg["attn_pos"])
        for i, v in enumerate(cfg["attn_pos"]):
            c.attn_pos[i] = v
        c.attn_with_skip, c.act_relu = int(bool(cfg["attn_with_skip"])), int(cfg["act_type"] == "relu")
        c.num_embeddings, c.num_groups = cfg["num_embeddings"], cfg["num_groups"]
        c.batch_invariant = int(bool(cfg.get("batch_invariant", False)))
        self.h = C.c_void_p()
        with torch.cuda.device(device):
            check(self.lib.ds_vqgan_create(C.byref(c), C.byref(self.h)), "ds_vqgan_create")
            skip = ("_vq_vae._ema_w", "_vq_vae._ema_cluster_size")          # training-time EMA state, unused by forward
            _load_params(self.lib.ds_vqgan_load, self.h, {k: v for k, v in state_dict.items() if k not in skip and "temb_proj" not in k})
            check(self.lib.ds_vqgan_finalize(self.h), "ds_vqgan_finalize")

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.ds_vqgan_destroy(h)
            except Exception:
                pass

    def decode(self, q: torch.Tensor) -> torch.Tensor:
        B, _, H, W = q.shape
        out = torch.empty((B, self.cfg["out_channels"], 4 * H, 4 * W), dtype=torch.float32, device=self.device)
        check(self.lib.ds_vqgan_decode(self.h, q.data_ptr(), out.data_ptr(), B, H, W, _stream()), "ds_vqgan_decode")
        return out

    def encode(self, spec: torch.Tensor) -> torch.Tensor:
        B, _, H, W = spec.shape
        out = torch.empty((B, self.cfg["embedding_dim"], H // 4, W // 4), dtype=torch.float32, device=self.device)
        check(self.lib.ds_vqgan_encode(self.h, spec.data_ptr(), out.data_ptr(), B, H, W, _stream()), "ds_vqgan_encode")
        return out


class SampleGraph:
    """ds_sample_graph handle: all steps of a sampling call (+ the VQ -> decoder -> iSTFT tail) as one CUDA graph over caller-owned
    device buffers (torch tensors held by the caller for the lifetime of this object)."""

    def __init__(self, unet: UnetEngine, vqgan: Optional[VqganEngine], bufs: SampleBuffers, B: int, H: int, W: int, n_iter: int,
                 cfg_on: bool, use_graph: bool = True):
        self.lib = _lib.load()
        self.unet, self.vqgan, self.bufs = unet, vqgan, bufs       # keep the handles this graph points into alive
        self.h = C.c_void_p()
        check(self.lib.ds_sample_graph_build(unet.h, vqgan.h if vqgan is not None else None, C.byref(bufs), B, H, W, n_iter, int(cfg_on),
                                             int(use_graph), _stream(), C.byref(self.h)), "ds_sample_graph_build")
        self.launches = int(self.lib.ds_sample_graph_launches(self.h))

    def run(self) -> None:
        check(self.lib.ds_sample_graph_run(self.h, _stream()), "ds_sample_graph_run")

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.ds_sample_graph_destroy(h)
            except Exception:
                pass


class Comm:
    """ds_comm handle: the job's one collective (all-gather of rank-local results) on the library's own NCCL communicator.
    ``bootstrap(buf)`` must broadcast the 128-byte id tensor from rank 0 to every rank (e.g. torch.distributed.broadcast on any
    backend): that exchange is the only thing the host contributes."""

    def __init__(self, rank: int, world: int, bootstrap=None):
        self.lib = _lib.load()
        self.rank, self.world = rank, world
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            check(self.lib.ds_comm_unique_id(ident.data_ptr()), "ds_comm_unique_id")
        if world > 1:
            assert bootstrap is not None, "a bootstrap broadcast is required for world > 1"
            ident = bootstrap(ident)
        self.h = C.c_void_p()
        check(self.lib.ds_comm_init(rank, world, ident.cpu().contiguous().data_ptr(), C.byref(self.h)), "ds_comm_init")

    def all_gather(self, send: torch.Tensor, recv: Optional[torch.Tensor] = None) -> torch.Tensor:
        send = send.contiguous()
        dt = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 1, torch.int64: 2}[send.dtype]
        if recv is None:
            recv = send.new_empty((self.world * send.shape[0],) + tuple(send.shape[1:]))
        assert recv.is_contiguous() and recv.numel() == self.world * send.numel() and recv.dtype == send.dtype
        check(self.lib.ds_allgather(self.h, send.data_ptr(), recv.data_ptr(), send.numel(), dt, _stream()), "ds_allgather")
        return recv

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.ds_comm_destroy(h)
            except Exception:
                pass

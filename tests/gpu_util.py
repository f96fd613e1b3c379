"""Helpers for the GPU parity tests (torch is used only to move/convert test data)."""
import torch

from diffusynth_b200 import ops

ACT = ops.ACT
# relative rounding step of the 16-bit operand type: fp16 2^-11, bf16 2^-8
EPS16 = 2.0 ** -11 if ACT == torch.float16 else 2.0 ** -8


def nhwc(x_nchw: torch.Tensor, cp: int = None) -> torch.Tensor:
    """fp32 NCHW (CPU) -> bf16 NHWC on cuda, channels zero-padded to cp."""
    x = x_nchw.permute(0, 2, 3, 1).contiguous()
    if cp is not None and cp > x.shape[-1]:
        x = torch.nn.functional.pad(x, (0, cp - x.shape[-1]))
    return x.to(ACT).cuda().contiguous()


def nchw(x_nhwc: torch.Tensor, c: int = None) -> torch.Tensor:
    x = x_nhwc.float().cpu()
    if c is not None:
        x = x[..., :c]
    return x.permute(0, 3, 1, 2).contiguous()


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf(x: torch.Tensor) -> torch.Tensor:
    return x.to(ACT).float()

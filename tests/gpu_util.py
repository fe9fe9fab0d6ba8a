"""Shared helpers for the -m gpu parity tests (all compute goes through the C ABI via moge_b200.capi)."""
import ctypes as C

import torch
import torch.nn.functional as F

from moge_b200 import capi


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def stream():
    return capi.current_stream()


def dt(dtype):
    return capi.torch_dtype_code(dtype)


def to_padded_nhwc(x_nchw: torch.Tensor, dtype) -> torch.Tensor:
    """(B,C,H,W) -> (B,Hp,Wp,C) with a replicated 1-pixel border; Hp = max(H+2, 8), Wp = max(W+2, 16)."""
    B, Cc, H, W = x_nchw.shape
    p = F.pad(x_nchw, (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1)
    Hp, Wp = max(H + 2, 8), max(W + 2, 16)
    out = torch.zeros(B, Hp, Wp, Cc, dtype=dtype, device=x_nchw.device)
    out[:, :H + 2, :W + 2] = p.to(dtype)
    return out.contiguous()


def empty_padded(B, H, W, Cc, dtype, device):
    return torch.full((B, max(H + 2, 8), max(W + 2, 16), Cc), float("nan"), dtype=dtype, device=device)


def from_padded_nhwc(x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """interior of a padded map -> (B,C,H,W) fp32"""
    return x[:, 1:H + 1, 1:W + 1].permute(0, 3, 1, 2).float()


def border_ok(x: torch.Tensor, H: int, W: int) -> bool:
    """the 1-pixel border must replicate the edge pixels"""
    inner = x[:, 1:H + 1, 1:W + 1].float()
    ref = F.pad(inner.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1)
    return bool(torch.equal(ref, x[:, :H + 2, :W + 2].float()))

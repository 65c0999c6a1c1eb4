"""Positional-embedding math of the reference (`opensora/models/mmdit/math.py:39-57`), host side only: the
rotation itself is applied inside the osb200 attention kernel while it stages q/k (no q/k round trip)."""
from __future__ import annotations

import torch
from torch import Tensor


def rope(pos: Tensor, dim: int, theta: int) -> Tensor:
    """math.py:50-57: fp64 angles -> [b, n, dim/2, 2, 2] rotation matrices in fp32."""
    assert dim % 2 == 0
    scale = torch.arange(0, dim, 2, dtype=torch.float64, device=pos.device) / dim
    omega = 1.0 / (theta**scale)
    out = torch.einsum("...n,d->...nd", pos.double(), omega)
    out = torch.stack([torch.cos(out), -torch.sin(out), torch.sin(out), torch.cos(out)], dim=-1)
    return out.reshape(*out.shape[:-1], 2, 2).float()


def liger_rope(pos: Tensor, dim: int, theta: int):
    """math.py:39-47."""
    assert dim % 2 == 0
    scale = torch.arange(0, dim, 2, dtype=torch.float32, device=pos.device) / dim
    omega = 1.0 / (theta**scale)
    out = torch.einsum("...n,d->...nd", pos, omega)
    return out.cos(), out.sin()


def rope_tables(pe) -> tuple[Tensor, Tensor, bool]:
    """(cos [L, D/2], sin [L, D/2], half_layout) fp32 tables for the attention kernel from either embedder's
    output.  Positions are shared by every sample of a batch in the reference's samplers
    (utils/sampling.py:437-447), which is what lets one table serve the batch; this is asserted."""
    if isinstance(pe, Tensor):  # EmbedND: [B, 1, L, D/2, 2, 2], interleaved pairs (2i, 2i+1)
        if pe.shape[0] > 1 and not torch.equal(pe[0], pe[-1]):
            raise NotImplementedError("osb200 attention takes one RoPE table per call: per-sample position ids differ")
        m = pe[0, 0]
        return m[..., 0, 0].contiguous(), m[..., 1, 0].contiguous(), False
    cos, sin = pe  # LigerEmbedND: [B, L, D] with the D/2 frequencies repeated (rotate-half pairs (i, i + D/2))
    if cos.shape[0] > 1 and not torch.equal(cos[0], cos[-1]):
        raise NotImplementedError("osb200 attention takes one RoPE table per call: per-sample position ids differ")
    d = cos.shape[-1] // 2
    return cos[0, :, :d].float().contiguous(), sin[0, :, :d].float().contiguous(), True

"""ORACLE — test infrastructure only.  fp32 restatement of the reference's rectified-flow sampler numerics
(`opensora/utils/sampling.py:120-133,159-226,295-393`).  PINNED by tests/golden/sampling.npz, produced by executing the
reference's own `sampling.py` (tests/golden/make_golden_sampling.py).  Only tests/ may import this module."""
from __future__ import annotations

import math

import torch


def time_shift(alpha, t):
    """sampling.py:295-296."""
    return alpha * t / (1 + (alpha - 1) * t)


def get_schedule(num_steps, image_seq_len, num_frames, shift_alpha=None, base_shift=1, max_shift=3, shift=True):
    """sampling.py:307-332."""
    ts = torch.linspace(1, 0, num_steps + 1)
    if shift:
        if shift_alpha is None:
            m = (max_shift - base_shift) / (4096 - 256)
            shift_alpha = (m * image_seq_len + (base_shift - m * 256)) * math.sqrt(num_frames)
        ts = time_shift(shift_alpha, ts)
    return ts.tolist()


def pack(x, p=2):
    """sampling.py:375-378."""
    b, c, t, hh, ww = x.shape
    return x.reshape(b, c, t, hh // p, p, ww // p, p).permute(0, 2, 3, 5, 1, 4, 6).reshape(b, t * (hh // p) * (ww // p), c * p * p)


def oscillation(gs, i, force_num=10):
    """sampling.py:120-133."""
    return gs if (i < force_num or i % 2 == 0) else 1.0


def denoise(model, img, timesteps, guidance, guidance_img, masks, masked_ref, text_osci=False, image_osci=False,
            scale_temporal_osci=False, patch_size=2, **kw):
    """sampling.py:159-226."""
    gvec = torch.full((img.shape[0],), guidance, dtype=img.dtype)
    for i, (tc, tp) in enumerate(zip(timesteps[:-1], timesteps[1:])):
        tvec = torch.full((img.shape[0],), tc, dtype=img.dtype)
        b, c, t, w, h = masked_ref.size()
        cond = pack(torch.cat((masks, masked_ref), dim=1), patch_size)
        x = img[: len(img) // 3]
        img = torch.cat([x, x, x], 0)
        pred = model(img=img, **kw, cond=torch.cat([cond, cond, torch.zeros_like(cond)], 0), timesteps=tvec, guidance=gvec)
        tg = oscillation(guidance, i) if text_osci else guidance
        ig = oscillation(guidance_img, i) if image_osci else guidance_img
        c_, u_, u2_ = pred.chunk(3, 0)
        if ig > 1.0 and scale_temporal_osci:
            upper = torch.linspace(ig, 1.0, len(timesteps))[i]
            ig = pack(torch.linspace(1.0, upper, t)[None, None, :, None, None].repeat(b, c, 1, h, w), patch_size).to(c_.dtype)
        pred = u2_ + ig * (u_ - u2_) + tg * (c_ - u_)
        img = img + (tp - tc) * torch.cat([pred, pred, pred], 0)
    return img[: len(img) // 3]

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


# ---- STDiT3's sampler (Open-Sora v1.2 `schedulers/rf`): NOT in the reference tree -> restatement of SURVEY.md Appendix A,
# ---- "RF sampler" row; PARITY UNPINNED (no reference source or vector exists for it here) --------------------------------
def rflow_timestep_transform(t, height, width, num_frames, num_timesteps=1000):
    """t' = r t / (1 + (r - 1) t),  r = sqrt(H W / 512^2) * sqrt(frames // 17 * 5)  (r_time = 1 for a single frame)."""
    t = t / num_timesteps
    r = math.sqrt(height * width / (512 * 512)) * (1.0 if num_frames == 1 else math.sqrt(num_frames // 17 * 5))
    return r * t / (1 + (r - 1) * t) * num_timesteps


def rflow_sample(model, z, y, y_null, mask=None, steps=30, cfg_scale=7.0, transform=None, **model_kw):
    """timesteps_i = (1 - i/N) 1000 (optionally transformed); per step z_in = cat[z, z], pred = model(z_in, cat[t, t],
    y = cat[y, y_null]).chunk(2, dim=1)[0]; v = v_u + s (v_c - v_u); z += v (t_i - t_{i+1}) / 1000."""
    B = z.shape[0]
    ts = [(1.0 - i / steps) * 1000.0 for i in range(steps)]
    if transform is not None:
        ts = [rflow_timestep_transform(t, *transform) for t in ts]
    kw = {k: (torch.cat((v, v), 0) if isinstance(v, torch.Tensor) and v.shape[:1] == (B,) else v) for k, v in model_kw.items()}
    if mask is not None:
        kw["mask"] = torch.cat((mask, mask), 0)
    for i, t in enumerate(ts):
        tv = torch.full((2 * B,), t, dtype=torch.float32)
        pred = model(torch.cat((z, z), 0), tv, y=torch.cat((y, y_null), 0), **kw).chunk(2, dim=1)[0]
        vc, vu = pred.chunk(2, dim=0)
        t_next = ts[i + 1] if i + 1 < len(ts) else 0.0
        z = z + (vu + cfg_scale * (vc - vu)) * ((t - t_next) / 1000.0)
    return z

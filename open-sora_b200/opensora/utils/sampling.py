"""Numeric core of the reference's rectified-flow sampler (`opensora/utils/sampling.py`) with the same function
names and argument meaning: schedule (`time_shift` :295, `get_res_lin_function` :299-304, `get_schedule` :307-332),
noise / packing (`get_noise` :335-372, `pack` / `unpack` :375-393), oscillating guidance (`get_oscillation_gs`
:120-133) and the denoising loop `I2VDenoiser.denoise` (:159-226), whose per-step CFG combine + Euler update runs
as ONE osb200 kernel (`osb_cfg_euler`) instead of ~7 elementwise torch ops.  Text encoding / prompt handling /
model construction (`prepare`, `prepare_api`, :401-726) are outside the hot path."""
from __future__ import annotations

import math
import os

import torch
from torch import Tensor


def time_shift(alpha: float, t: Tensor) -> Tensor:
    return alpha * t / (1 + (alpha - 1) * t)


def get_res_lin_function(x1: float = 256, y1: float = 1, x2: float = 4096, y2: float = 3):
    m = (y2 - y1) / (x2 - x1)
    b = y1 - m * x1
    return lambda x: m * x + b


def get_schedule(num_steps: int, image_seq_len: int, num_frames: int, shift_alpha: float | None = None,
                 base_shift: float = 1, max_shift: float = 3, shift: bool = True) -> list[float]:
    timesteps = torch.linspace(1, 0, num_steps + 1)
    if shift:
        if shift_alpha is None:
            shift_alpha = get_res_lin_function(y1=base_shift, y2=max_shift)(image_seq_len)
            shift_alpha *= math.sqrt(num_frames)
        timesteps = time_shift(shift_alpha, timesteps)
    return timesteps.tolist()


def get_noise(num_samples: int, height: int, width: int, num_frames: int, device, dtype, seed: int, patch_size: int = 2,
              channel: int = 16) -> Tensor:
    D = int(os.environ.get("AE_SPATIAL_COMPRESSION", 16))
    return torch.randn(num_samples, channel, num_frames, patch_size * math.ceil(height / D), patch_size * math.ceil(width / D),
                       device=device, dtype=dtype, generator=torch.Generator(device=device).manual_seed(seed))


def pack(x: Tensor, patch_size: int = 2) -> Tensor:
    """"b c t (h ph) (w pw) -> b (t h w) (c ph pw)" (:375-378)."""
    b, c, t, hh, ww = x.shape
    h, w = hh // patch_size, ww // patch_size
    x = x.reshape(b, c, t, h, patch_size, w, patch_size).permute(0, 2, 3, 5, 1, 4, 6)
    return x.reshape(b, t * h * w, c * patch_size * patch_size)


def unpack(x: Tensor, height: int, width: int, num_frames: int, patch_size: int = 2) -> Tensor:
    """"b (t h w) (c ph pw) -> b c t (h ph) (w pw)" (:381-393)."""
    D = int(os.environ.get("AE_SPATIAL_COMPRESSION", 16))
    h, w, t = math.ceil(height / D), math.ceil(width / D), num_frames
    b, _, cpp = x.shape
    c = cpp // (patch_size * patch_size)
    x = x.reshape(b, t, h, w, c, patch_size, patch_size).permute(0, 4, 1, 2, 5, 3, 6)
    return x.reshape(b, c, t, h * patch_size, w * patch_size)


def get_oscillation_gs(guidance_scale: float, i: int, force_num=10):
    if i < force_num or (i >= force_num and i % 2 == 0):
        return guidance_scale
    return 1.0


class I2VDenoiser:
    """`I2VDenoiser.denoise` (:159-226): 3-way CFG batch (cond / uncond-text / uncond-text+image), Euler steps."""

    def denoise(self, model, **kwargs) -> Tensor:
        import osb200

        img = kwargs.pop("img")
        timesteps = kwargs.pop("timesteps")
        guidance = kwargs.pop("guidance")
        guidance_img = kwargs.pop("guidance_img")
        masks = kwargs.pop("masks")
        masked_ref = kwargs.pop("masked_ref")
        kwargs.pop("sigma_min", None)
        text_osci = kwargs.pop("text_osci", False)
        image_osci = kwargs.pop("image_osci", False)
        scale_temporal_osci = kwargs.pop("scale_temporal_osci", False)
        patch_size = kwargs.pop("patch_size", 2)

        guidance_vec = torch.full((img.shape[0],), guidance, device=img.device, dtype=img.dtype)
        b, c, t, w, h = masked_ref.size()  # (sic) the reference names them this way (:186)
        cond = pack(torch.cat((masks, masked_ref), dim=1), patch_size=patch_size)
        kwargs["cond"] = torch.cat([cond, cond, torch.zeros_like(cond)], dim=0)  # step-invariant: hoisted out of the loop
        x = img[: len(img) // 3].contiguous()
        for i, (t_curr, t_prev) in enumerate(zip(timesteps[:-1], timesteps[1:])):
            t_vec = torch.full((img.shape[0],), t_curr, dtype=img.dtype, device=img.device)
            pred = model(img=torch.cat([x, x, x], dim=0), **kwargs, timesteps=t_vec, guidance=guidance_vec)
            text_gs = get_oscillation_gs(guidance, i) if text_osci else guidance
            image_gs = get_oscillation_gs(guidance_img, i) if image_osci else guidance_img
            gmap = None
            if image_gs > 1.0 and scale_temporal_osci:
                upper = torch.linspace(image_gs, 1.0, len(timesteps))[i]
                g5 = torch.linspace(1.0, upper, t)[None, None, :, None, None].repeat(b, c, 1, h, w)
                gmap = pack(g5, patch_size=patch_size).to(pred.device, pred.dtype).contiguous()
                image_gs = 1.0
            pc, pu, pu2 = (p.contiguous() for p in pred.chunk(3, dim=0))
            # pred = uncond_2 + image_gs*(uncond - uncond_2) + text_gs*(cond - uncond); x += (t_prev - t_curr)*pred  (:219-222)
            x = osb200.cfg_euler(pc, pu, pu2, x, g_txt=float(text_gs), g_img=float(image_gs), g_img_map=gmap,
                                 dt=float(t_prev - t_curr))
        return x

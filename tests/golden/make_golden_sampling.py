"""Generates tests/golden/sampling.npz by EXECUTING the reference's own `opensora/utils/sampling.py` (loaded by path,
oracle/ref_loader.load_sampling) with a deterministic toy denoiser.  Build container only."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402


def toy_model(img, cond, timesteps, guidance, **kw):
    """Deterministic stand-in for the denoiser: depends on every input the loop feeds it."""
    scale = torch.tensor([1.0, 0.5, 0.25]).repeat_interleave(img.shape[0] // 3)[:, None, None]
    return torch.tanh(img * 0.7 + cond[..., : img.shape[-1]] * 0.3) * scale * (1 + timesteps[:, None, None]) + 0.01 * guidance[:, None, None]


def main():
    S = ref_loader.load_sampling()
    out = {}
    out["sched_50_shift"] = np.array(S.get_schedule(50, 12 * 21, 33), dtype=np.float64)
    out["sched_8_noshift"] = np.array(S.get_schedule(8, 64, 5, shift=False), dtype=np.float64)
    out["sched_10_alpha3"] = np.array(S.get_schedule(10, 64, 5, shift_alpha=3.0), dtype=np.float64)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(2, 16, 3, 8, 12, generator=g)
    out["pack_in"] = z.numpy()
    out["pack_out"] = S.pack(z).numpy()
    os.environ["AE_SPATIAL_COMPRESSION"] = "16"
    out["unpack_out"] = S.unpack(S.pack(z), 64, 96, 3).numpy()
    out["osc"] = np.array([S.get_oscillation_gs(7.5, i) for i in range(14)])
    B = 2
    img = S.pack(z).repeat(3, 1, 1)
    masks = torch.zeros(B, 1, 3, 8, 12)
    masks[:, :, 0] = 1
    masked_ref = torch.randn(B, 16, 3, 8, 12, generator=g) * masks
    ts = S.get_schedule(12, 24, 3)
    for tag, kw in (("plain", {}), ("osci", dict(text_osci=True, image_osci=True, scale_temporal_osci=True))):
        res = S.I2VDenoiser().denoise(toy_model, img=img.clone(), timesteps=ts, guidance=7.5, guidance_img=3.0, masks=masks,
                                      masked_ref=masked_ref, sigma_min=1e-5, patch_size=2, **kw)
        out[f"denoise_{tag}"] = res.numpy()
    out.update(den_img=img.numpy(), den_masks=masks.numpy(), den_ref=masked_ref.numpy(), den_ts=np.array(ts))
    path = os.path.join(HERE, "sampling.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

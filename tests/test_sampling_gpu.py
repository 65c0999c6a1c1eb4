"""GPU: the fused CFG + Euler kernel and the osb200 `I2VDenoiser.denoise` loop against the reference-executed golden."""
import os

import numpy as np
import pytest
import torch

from tests.util import report

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _toy(img, cond, timesteps, guidance, **kw):
    scale = torch.tensor([1.0, 0.5, 0.25], device=img.device).repeat_interleave(img.shape[0] // 3)[:, None, None]
    r = torch.tanh(img.float() * 0.7 + cond[..., : img.shape[-1]].float() * 0.3) * scale * (1 + timesteps.float()[:, None, None])
    return (r + 0.01 * guidance.float()[:, None, None]).to(img.dtype)


def test_cfg_euler_kernel():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import osb200

    g = torch.Generator(device="cuda").manual_seed(0)
    c, u, u2, x = (torch.randn(3, 1000, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
    gmap = (1 + torch.rand(1, 1000, 64, device="cuda", generator=g)).to(torch.bfloat16)
    for kw, ref in (
        (dict(g_txt=7.5, g_img=3.0, dt=-0.02), x.float() - 0.02 * (u2.float() + 3.0 * (u.float() - u2.float()) + 7.5 * (c.float() - u.float()))),
        (dict(g_txt=7.5, g_img=1.0, g_img_map=gmap, dt=-0.02), x.float() - 0.02 * (u2.float() + gmap.float() * (u.float() - u2.float()) + 7.5 * (c.float() - u.float()))),
    ):
        out = osb200.cfg_euler(c, u, u2, x, **kw)
        r, _ = report("cfg_euler", out, ref)
        assert r < 2e-3
    out = osb200.cfg_euler(c, u, None, x, g_txt=7.0, dt=-0.1)
    ref = x.float() - 0.1 * (u.float() + 7.0 * (c.float() - u.float()))
    assert report("cfg_euler 2-way", out, ref)[0] < 2e-3


@pytest.mark.parametrize("tag,kw", [("plain", {}), ("osci", dict(text_osci=True, image_osci=True, scale_temporal_osci=True))])
def test_denoise_loop_vs_reference_golden(tag, kw):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from opensora.utils.sampling import I2VDenoiser

    G = dict(np.load(os.path.join(HERE, "golden", "sampling.npz")))
    cu = lambda k: torch.from_numpy(G[k]).cuda()  # noqa: E731
    out = I2VDenoiser().denoise(_toy, img=cu("den_img").to(torch.bfloat16), timesteps=list(G["den_ts"]), guidance=7.5,
                                guidance_img=3.0, masks=cu("den_masks").to(torch.bfloat16), masked_ref=cu("den_ref").to(torch.bfloat16),
                                sigma_min=1e-5, patch_size=2, **kw)
    r, _ = report(f"denoise loop {tag} (12 steps, bf16) vs reference fp32 golden", out, cu(f"denoise_{tag}"))
    assert r < 2e-2

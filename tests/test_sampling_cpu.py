"""Sampler numerics: the host-side mirror `opensora.utils.sampling` and the oracle restatement
against fixtures produced by EXECUTING the reference's own `opensora/utils/sampling.py`
(tests/golden/make_golden_sampling.py).  The fused CFG+Euler kernel itself is checked on the GPU (test_sampling_gpu)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
G = dict(np.load(os.path.join(HERE, "golden", "sampling.npz")))


def _toy(img, cond, timesteps, guidance, **kw):
    scale = torch.tensor([1.0, 0.5, 0.25]).repeat_interleave(img.shape[0] // 3)[:, None, None]
    return torch.tanh(img * 0.7 + cond[..., : img.shape[-1]] * 0.3) * scale * (1 + timesteps[:, None, None]) + 0.01 * guidance[:, None, None]


def test_schedule_pack_unpack_match_reference():
    from opensora.utils import sampling as S
    from oracle import sampling_oracle as O

    for mod in (S, O):
        np.testing.assert_allclose(mod.get_schedule(50, 12 * 21, 33), G["sched_50_shift"], rtol=1e-6)
        np.testing.assert_allclose(mod.get_schedule(8, 64, 5, shift=False), G["sched_8_noshift"], rtol=1e-6)
        np.testing.assert_allclose(mod.get_schedule(10, 64, 5, shift_alpha=3.0), G["sched_10_alpha3"], rtol=1e-6)
        assert np.array_equal(mod.pack(torch.from_numpy(G["pack_in"])).numpy(), G["pack_out"])
    assert np.array_equal(S.unpack(torch.from_numpy(G["pack_out"]), 64, 96, 3).numpy(), G["unpack_out"])
    assert [S.get_oscillation_gs(7.5, i) for i in range(14)] == list(G["osc"])
    assert [O.oscillation(7.5, i) for i in range(14)] == list(G["osc"])


def test_oracle_denoise_loop_matches_reference():
    from oracle import sampling_oracle as O

    for tag, kw in (("plain", {}), ("osci", dict(text_osci=True, image_osci=True, scale_temporal_osci=True))):
        out = O.denoise(_toy, torch.from_numpy(G["den_img"]), list(G["den_ts"]), 7.5, 3.0, torch.from_numpy(G["den_masks"]),
                        torch.from_numpy(G["den_ref"]), **kw)
        np.testing.assert_allclose(out.numpy(), G[f"denoise_{tag}"], rtol=1e-5, atol=1e-5)


# ---- the denoise loop itself (host logic of opensora/utils/sampling.py::I2VDenoiser) through the CPU stand-in of the
# ---- binding, against the golden produced by executing the reference's sampling.py ----------------------------------
def _toy(img, cond, timesteps, guidance, **kw):
    scale = torch.tensor([1.0, 0.5, 0.25]).repeat_interleave(img.shape[0] // 3)[:, None, None]
    r = torch.tanh(img.float() * 0.7 + cond[..., : img.shape[-1]].float() * 0.3) * scale * (1 + timesteps.float()[:, None, None])
    return (r + 0.01 * guidance.float()[:, None, None]).to(img.dtype)


@pytest.mark.parametrize("tag,kw", [("plain", {}), ("osci", dict(text_osci=True, image_osci=True, scale_temporal_osci=True))])
def test_denoise_loop_host_logic_vs_reference_golden(fake_osb, tag, kw):
    import os

    import numpy as np

    from opensora.utils.sampling import I2VDenoiser

    G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampling.npz")))
    t = lambda k: torch.from_numpy(G[k])  # noqa: E731
    out = I2VDenoiser().denoise(_toy, img=t("den_img").to(torch.bfloat16), timesteps=list(G["den_ts"]), guidance=7.5,
                                guidance_img=3.0, masks=t("den_masks").to(torch.bfloat16), masked_ref=t("den_ref").to(torch.bfloat16),
                                sigma_min=1e-5, patch_size=2, **kw)
    ref = t(f"denoise_{tag}")
    r = float((out.float() - ref).norm() / ref.norm())
    assert r < 2e-2, r
    assert [c[0] for c in fake_osb.calls].count("cfg_euler") == len(G["den_ts"]) - 1   # one fused update per step

"""Sampler numerics: the host-side mirror `opensora.utils.sampling` and the oracle restatement
against fixtures produced by EXECUTING the reference's own `opensora/utils/sampling.py`
(tests/golden/make_golden_sampling.py).  The fused CFG+Euler kernel itself is checked on the GPU (test_sampling_gpu)."""
import os

import numpy as np
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

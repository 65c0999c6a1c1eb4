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


# ---- the request side of the sampler (options, prompts, model inputs, conditioning, prepare_api) against goldens produced by
# ---- executing the reference's sampling.py + inference.py + datasets/aspect.py (make_golden_sampling.py::pipeline) ---------
def _pipeline_golden():
    import json

    g = dict(np.load(os.path.join(HERE, "golden", "sampling_pipeline.npz")))
    return g, json.loads(str(g["meta"]))


def test_options_aspect_tables_and_guidance_prompts():
    from opensora.datasets import aspect as A
    from opensora.utils import inference as I
    from opensora.utils import sampling as S

    _, meta = _pipeline_golden()
    for res, table in meta["aspect_inference"].items():
        mine = A.get_aspect_ratios_dict(A.get_num_pexels_from_name(res), training=False)
        assert {k: list(v) for k, v in mine.items()} == table and list(mine) == list(table), res
    for res, table in meta["aspect_training"].items():
        mine = A.get_aspect_ratios_dict(A.get_num_pexels_from_name(res), training=True)
        assert {k: list(v) for k, v in mine.items()} == table and list(mine) == list(table), res
    cases = dict(res=dict(resolution="768px", aspect_ratio="9:16", method="i2v"), hw=dict(height=250, width=443),
                 hw16=dict(height=256, width=448, method="distill"), res360=dict(resolution="360p", aspect_ratio="2.39:1"))
    for name, kw in cases.items():
        o = S.sanitize_sampling_option(S.SamplingOption(**kw))
        assert [o.height, o.width, o.method.value] == meta["sanitize"][name], name
    with pytest.raises(AssertionError):
        S.sanitize_sampling_option(S.SamplingOption(resolution="256px"))           # needs the aspect ratio too
    with pytest.raises(AssertionError):
        S.sanitize_sampling_option(S.SamplingOption(height=64))                    # needs the width too
    with pytest.raises(ValueError):
        A.get_num_pexels_from_name("big")
    t2i = I.modify_option_to_t2i(S.SamplingOption(resolution="256px", aspect_ratio="16:9", num_frames=33, guidance=7.5), distilled=True,
                                 img_resolution="768px")
    assert [t2i.height, t2i.width, t2i.num_frames, t2i.guidance, t2i.method.value, t2i.resized_resolution] == meta["t2i"]
    assert S.I2VDenoiser().prepare_guidance(["a", "b"], {}, "cpu", torch.float32, neg=None, guidance_img=3.0) == \
        (meta["guidance_i2v"], {"guidance_img": 3.0})
    assert S.I2VDenoiser().prepare_guidance(["a"], {}, "cpu", torch.float32, neg=["n"], guidance_img=3.0)[0] == meta["guidance_i2v_neg"]
    assert S.DistilledDenoiser().prepare_guidance(["a"], {}, "cpu", torch.float32) == (["a"], {})
    assert set(S.SamplingMethodDict) == {S.SamplingMethod.I2V, S.SamplingMethod.DISTILLED}
    # prompt suffix conventions (fps / motion score) of the request format
    assert I.add_fps_info_to_text(list(meta["prompts"]), fps=24) == meta["fps_text"]
    assert I.add_fps_info_to_text(list(meta["prompts"])) == meta["fps_text_default"]
    assert I.add_motion_score_to_text(list(meta["prompts"]), 4) == meta["motion_text"]
    assert I.add_motion_score_to_text(["a"], "dynamic", refine_prompts=lambda t, type: ["7 motion score"]) == ["a 7 motion score."]
    with pytest.raises(NotImplementedError):
        I.add_motion_score_to_text(["a"], "dynamic")


def test_model_inputs_and_conditioning_format():
    from opensora.utils import inference as I
    from opensora.utils import sampling as S
    from tests import sampling_toys as T

    g, _ = _pipeline_golden()
    z = torch.from_numpy(g["prepare_z"])
    mine = S.prepare(T.toy_t5, T.toy_clip, z, prompt=["a cat", "neg", "neg"])
    ids = S.prepare_ids(z.repeat(2, 1, 1, 1, 1), torch.from_numpy(g["ids_t5"]), torch.from_numpy(g["ids_clip"]))
    for tag, d in (("prepare", mine), ("prepare_ids", ids)):
        assert sorted(d) == sorted(k.split(".", 1)[1] for k in g if k.startswith(tag + "."))
        for k, v in d.items():
            assert v.shape == g[f"{tag}.{k}"].shape and np.array_equal(v.numpy(), g[f"{tag}.{k}"]), (tag, k)
    refs = [list(torch.from_numpy(g["cond_refs"])), None]
    zc = torch.zeros(2, 4, 20, 2, 3)
    for kind in ("t2v", "i2v_head", "i2v_tail", "i2v_loop", "v2v_head", "v2v_tail", "v2v_head_easy", "v2v_tail_easy"):
        for causal in (True, False):
            m, mz = I.prepare_inference_condition(zc, kind, ref_list=refs, causal=causal)
            assert np.array_equal(m.numpy(), g[f"cond.{kind}.{int(causal)}.masks"]), (kind, causal)
            assert np.array_equal(mz.numpy(), g[f"cond.{kind}.{int(causal)}.ref"]), (kind, causal)
    with pytest.raises(AssertionError):
        I.prepare_inference_condition(zc, "i2v_middle", ref_list=refs)
    m, mz = I.prepare_inference_condition(torch.zeros(1, 4, 1, 2, 3), "i2v_head", ref_list=[refs[0]])   # an image: nothing pinned
    assert not m.any() and not mz.any()


def test_distilled_denoiser_matches_reference():
    from opensora.utils import sampling as S
    from tests import sampling_toys as T

    g, _ = _pipeline_golden()
    out = S.DistilledDenoiser().denoise(T.ToyDenoiser(), img=torch.from_numpy(g["distill_x0"]), timesteps=S.get_schedule(5, 24, 1),
                                        guidance=3.5, img_ids=torch.zeros(2, 24, 3), txt=torch.ones(2, 6, 8),
                                        txt_ids=torch.zeros(2, 6, 3), y_vec=torch.ones(2, 8))
    np.testing.assert_allclose(out.detach().numpy(), g["distill_out"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("scenario", [s[0] for s in __import__("tests.sampling_toys", fromlist=["SCENARIOS"]).SCENARIOS])
def test_prepare_api_end_to_end_vs_reference(fake_osb, monkeypatch, scenario):
    """`prepare_api(...)(opt, cond_type, text=..., ref=...)` - noise, schedule, prompts, model inputs, conditioning, the
    denoising loop, reference re-insertion, decode and the frame crops - on the toy models, against the reference's own
    `prepare_api` executed on the same toys.  The i2v loop runs in bf16 through the binding stand-in (the fused CFG + Euler
    kernel's contract), the golden in fp32: rel-L2 bar 2e-2; the distilled (pure torch) case runs in fp32 and must agree to
    1e-3 (fp16 storage of the golden)."""
    from opensora.utils import sampling as S
    from tests import sampling_toys as T

    g, meta = _pipeline_golden()
    name, opt_kw, call_kw = next(s for s in T.SCENARIOS if s[0] == scenario)
    distilled = opt_kw.get("method") == "distill"
    dtype = torch.float32 if distilled else torch.bfloat16
    model = T.ToyDenoiser().to(dtype)
    ae = T.ToyAE(causal=opt_kw.get("is_causal_vae", False)).to(dtype)
    real_noise = S.get_noise
    # the CPU generator draws bf16 normals differently from fp32 ones; the golden run was fp32: draw fp32, round once
    monkeypatch.setattr(S, "get_noise", lambda n, h, w, f, device, dt, seed, **kw: real_noise(n, h, w, f, device, torch.float32, seed, **kw).to(dt))
    media = T.reference_media()
    api = S.prepare_api(model, ae, T.toy_t5, T.toy_clip, {})
    opt = S.sanitize_sampling_option(S.SamplingOption(**opt_kw))
    kw = {k: (list(v) if isinstance(v, list) else v) for k, v in call_kw.items()}
    if "ref" in kw:
        kw["reader"] = lambda path, image_size, transform_name=None: media[path]
    x = api(opt, **kw)
    ref = torch.from_numpy(g[f"api.{name}"].astype(np.float32))
    assert list(x.shape) == meta[f"api.{name}.shape"]
    assert len(model.seen) == meta[f"api.{name}.calls"] and model.seen[0] == meta[f"api.{name}.model_kwargs"]
    r = float((x.float() - ref).norm() / ref.norm())
    assert r < (1e-3 if distilled else 2e-2), (name, r)
    if not distilled:
        assert [c[0] for c in fake_osb.calls].count("cfg_euler") == opt.num_steps   # one fused update per step


# ---- STDiT3's own sampler (v1.2 RFLOW; absent from the reference tree, restated from SURVEY.md Appendix A: unpinned) ------------
def _toy_stdit(x, timestep, y, mask=None, fps=None, height=None, width=None, **kw):
    """[2B, C, T, H, W] -> [2B, 2C, T, H, W] fp32 (velocity | sigma halves), depending on every input."""
    f = x.float()
    cap = (y.float() * (1.0 if mask is None else mask.float()[:, None, :, None])).mean(dim=(1, 2, 3))[:, None, None, None, None]
    v = torch.tanh(0.8 * f + cap) * (0.5 + timestep.float()[:, None, None, None, None] / 1000.0) + 0.01 * fps.float()[:, None, None, None, None]
    return torch.cat((v, 0.1 * f), dim=1)


@pytest.mark.parametrize("transform", [False, True])
def test_rflow_sampler_host_logic_vs_oracle(fake_osb, transform):
    from opensora.registry import SCHEDULERS
    from opensora.schedulers import RFLOW, timestep_transform
    from oracle import sampling_oracle as O

    assert SCHEDULERS.get("rflow") is RFLOW
    g = torch.Generator().manual_seed(9)
    B, C, T, H, W, L = 2, 4, 5, 6, 8, 7
    z = torch.randn(B, C, T, H, W, generator=g)
    y, y_null = torch.randn(B, 1, L, 16, generator=g), torch.randn(1, 1, L, 16, generator=g).repeat(B, 1, 1, 1)
    mask = torch.ones(B, L)
    mask[1, 4:] = 0
    extra = dict(fps=torch.full((B,), 24.0), height=torch.full((B,), 360.0), width=torch.full((B,), 640.0))
    ref = O.rflow_sample(_toy_stdit, z.to(torch.bfloat16).float(), y, y_null, mask=mask, steps=8, cfg_scale=7.0,
                         transform=(360.0, 640.0, 51) if transform else None, **extra)
    sch = RFLOW(num_sampling_steps=8, cfg_scale=7.0, use_timestep_transform=transform)
    out = sch.sample(_toy_stdit, z.to(torch.bfloat16), y, y_null, mask=mask, additional_args=dict(extra, num_frames=torch.full((B,), 51)))
    r = float((out.float() - ref).norm() / ref.norm())
    assert out.dtype == torch.bfloat16 and r < 2e-2, r
    assert [c[0] for c in fake_osb.calls].count("cfg_euler") == 8          # one fused combine + Euler update per step
    # the schedule transform: identity at r = 1 (512 x 512 image), fixed points 0 and 1000, closed form at a known point
    t = torch.tensor([0.0, 250.0, 1000.0])
    assert torch.allclose(timestep_transform(t, 512.0, 512.0, torch.tensor([1])), t)
    tt = timestep_transform(t, 360.0, 640.0, torch.tensor([51]))
    r_ = (360 * 640 / 512**2) ** 0.5 * (51 // 17 * 5) ** 0.5
    assert torch.allclose(tt, torch.tensor([0.0, 1000 * r_ * 0.25 / (1 + (r_ - 1) * 0.25), 1000.0]), rtol=1e-5)
    assert abs(float(tt[1]) - O.rflow_timestep_transform(250.0, 360.0, 640.0, 51)) < 1e-3

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
    pipeline(S)


def pipeline(S):
    """tests/golden/sampling_pipeline.npz: the request-side half of the reference file, executed - option sanitising (with
    the reference's own datasets/aspect.py), guidance prompts, `prepare` / `prepare_ids`, the conditioning format
    (utils/inference.py), DistilledDenoiser and `prepare_api` end to end on the toy models of tests/sampling_toys.py."""
    import json

    from tests import sampling_toys as T

    out, meta = {}, {}
    A, I = S.ref_aspect, S.ref_inference
    meta["aspect_inference"] = {res: {k: list(v) for k, v in A.get_aspect_ratios_dict(A.get_num_pexels_from_name(res), False).items()}
                                for res in ("256px", "768px", "360p", "720p", "1080px")}
    meta["aspect_training"] = {res: {k: list(v) for k, v in A.get_aspect_ratios_dict(A.get_num_pexels_from_name(res), True).items()}
                               for res in ("256px", "768px")}
    san = {}
    for name, kw in (("res", dict(resolution="768px", aspect_ratio="9:16", method="i2v")), ("hw", dict(height=250, width=443)),
                     ("hw16", dict(height=256, width=448, method="distill")), ("res360", dict(resolution="360p", aspect_ratio="2.39:1"))):
        o = S.sanitize_sampling_option(S.SamplingOption(**kw))
        san[name] = [o.height, o.width, o.method.value]
    meta["sanitize"] = san
    t2i = I.modify_option_to_t2i(S.SamplingOption(resolution="256px", aspect_ratio="16:9", num_frames=33, guidance=7.5), distilled=True,
                                 img_resolution="768px")
    meta["t2i"] = [t2i.height, t2i.width, t2i.num_frames, t2i.guidance, t2i.method.value, t2i.resized_resolution]
    prompts = ["a cat", "a dog runs.  ", "waves at 24 FPS.", "x 16 FPS"]
    meta["prompts"] = prompts
    meta["fps_text"] = I.add_fps_info_to_text(list(prompts), fps=24)
    meta["fps_text_default"] = I.add_fps_info_to_text(list(prompts))
    meta["motion_text"] = I.add_motion_score_to_text(list(prompts), 4)
    meta["guidance_i2v"] = S.I2VDenoiser().prepare_guidance(["a", "b"], {}, "cpu", torch.float32, neg=None, guidance_img=3.0)[0]
    meta["guidance_i2v_neg"] = S.I2VDenoiser().prepare_guidance(["a"], {}, "cpu", torch.float32, neg=["n"], guidance_img=3.0)[0]

    g = torch.Generator().manual_seed(21)
    z = torch.randn(1, 4, 3, 8, 12, generator=g)
    for tag, d in (("prepare", S.prepare(T.toy_t5, T.toy_clip, z, prompt=["a cat", "neg", "neg"])),
                   ("prepare_ids", S.prepare_ids(z.repeat(2, 1, 1, 1, 1), torch.randn(1, 5, 8, generator=g), torch.randn(1, 8, generator=g)))):
        for k, v in d.items():
            out[f"{tag}.{k}"] = v.numpy()
    out["prepare_z"] = z.numpy()
    g = torch.Generator().manual_seed(21)
    torch.randn(1, 4, 3, 8, 12, generator=g)
    out["ids_t5"], out["ids_clip"] = torch.randn(1, 5, 8, generator=g).numpy(), torch.randn(1, 8, generator=g).numpy()

    # conditioning format, every kind, causal and not
    zc = torch.zeros(2, 4, 20, 2, 3)
    refs = [[torch.randn(4, 20, 2, 3, generator=g), torch.randn(4, 20, 2, 3, generator=g)], None]
    out["cond_refs"] = torch.stack(refs[0]).numpy()
    for kind in ("t2v", "i2v_head", "i2v_tail", "i2v_loop", "v2v_head", "v2v_tail", "v2v_head_easy", "v2v_tail_easy"):
        for causal in (True, False):
            m, mz = I.prepare_inference_condition(zc, kind, ref_list=refs, causal=causal)
            out[f"cond.{kind}.{int(causal)}.masks"], out[f"cond.{kind}.{int(causal)}.ref"] = m.numpy(), mz.numpy()

    model = T.ToyDenoiser()
    x0 = torch.randn(2, 24, 16, generator=g)
    out["distill_x0"] = x0.numpy()
    out["distill_out"] = S.DistilledDenoiser().denoise(
        model, img=x0.clone(), timesteps=S.get_schedule(5, 24, 1), guidance=3.5, img_ids=torch.zeros(2, 24, 3), txt=torch.ones(2, 6, 8),
        txt_ids=torch.zeros(2, 6, 3), y_vec=torch.ones(2, 8)).detach().numpy()

    media = T.reference_media()
    I.read_from_path = lambda path, image_size, transform_name=None: media[path]
    for name, opt_kw, call_kw in T.SCENARIOS:
        ae = T.ToyAE(causal=opt_kw.get("is_causal_vae", False))
        api = S.prepare_api(model, ae, T.toy_t5, T.toy_clip, {})
        opt = S.sanitize_sampling_option(S.SamplingOption(**opt_kw))
        model.seen.clear()
        x = api(opt, **{k: (list(v) if isinstance(v, list) else v) for k, v in call_kw.items()})
        out[f"api.{name}"] = x.numpy().astype(np.float16)
        meta[f"api.{name}.shape"] = list(x.shape)
        meta[f"api.{name}.model_kwargs"] = model.seen[0]
        meta[f"api.{name}.calls"] = len(model.seen)
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, "sampling_pipeline.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

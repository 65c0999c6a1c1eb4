"""The tensor-producing half of `opensora/utils/inference.py` that sits between the user's request and the denoiser:
`SamplingMethod` (:16-18), `modify_option_to_t2i` (:43-55), `add_noise_to_ref` (:210-213), `collect_references_batch`
(:216-280) and `prepare_inference_condition` (:283-351) - the image / video conditioning format (`masks`, `masked_ref`)
that `I2VDenoiser.denoise` packs into `cond` - plus the prompt suffix conventions the text conditioning was trained with
(`add_fps_info_to_text` :186-196, `add_motion_score_to_text` :199-207).  CSV handling, file naming, saving and LLM prompt
refinement (:21-40, :58-163) are I/O around the pipeline and out of scope."""
from __future__ import annotations

import copy
import re
from enum import Enum

import torch
from torch import nn

from opensora.datasets.aspect import get_image_size


class SamplingMethod(Enum):
    I2V = "i2v"            # Open-Sora video generation
    DISTILLED = "distill"  # guidance-distilled Flux image generation


def modify_option_to_t2i(sampling_option, distilled: bool = False, img_resolution: str = "1080px"):
    """The text-to-image pre-pass of t2i2v: one frame at `img_resolution`, same aspect ratio, guidance 4."""
    opt = copy.copy(sampling_option)
    if distilled:
        opt.method = SamplingMethod.DISTILLED
    opt.num_frames = 1
    opt.height, opt.width = get_image_size(img_resolution, sampling_option.aspect_ratio)
    opt.guidance = 4.0
    opt.resized_resolution = sampling_option.resolution
    return opt


def check_fps_added(sentence: str) -> bool:
    """Does the prompt already end with "<n> FPS."?"""
    return re.search(r"\d+ FPS\.$", sentence) is not None


def ensure_sentence_ends_with_period(sentence: str) -> str:
    sentence = sentence.strip()
    return sentence if sentence.endswith(".") else sentence + "."


def add_fps_info_to_text(text: list[str], fps: int = 16) -> list[str]:
    """Every prompt ends with a period and then " <fps> FPS." (once)."""
    out = []
    for prompt in text:
        prompt = ensure_sentence_ends_with_period(prompt)
        out.append(prompt if check_fps_added(prompt) else f"{prompt} {fps} FPS.")
    return out


def add_motion_score_to_text(text: list[str], motion_score: int | str, refine_prompts=None) -> list[str]:
    """Append "<score> motion score." to every prompt.  "dynamic" asks an LLM for a per-prompt score upstream
    (`opensora/utils/prompt_refine.py`, network access): pass that callable as `refine_prompts` to get the same behaviour."""
    if motion_score == "dynamic":
        if refine_prompts is None:
            raise NotImplementedError('motion_score="dynamic" needs `refine_prompts(text, type="motion_score")` (an LLM call upstream)')
        scores = refine_prompts(text, type="motion_score")
        return [f"{t} {scores[i]}." for i, t in enumerate(text)]
    return [f"{t} {motion_score} motion score." for t in text]


def add_noise_to_ref(masked_ref: torch.Tensor, masks: torch.Tensor, t: float, sigma_min: float = 1e-5):
    noise = torch.randn_like(masked_ref)
    return masks * ((1 - (1 - sigma_min) * t) * masked_ref + t * noise)


def _default_reader(path, image_size, transform_name="resize_crop"):
    """Reference media arrive as pixel tensors [C, T, H, W] (or `.pt` files holding one).  Decoding images / videos from
    disk (`opensora/datasets/utils.py::read_from_path`: PIL / av + resize-crop) is dataset I/O and not mirrored - pass
    `reader=` to plug one in."""
    if isinstance(path, torch.Tensor):
        return path
    if isinstance(path, str) and path.endswith(".pt"):
        return torch.load(path, map_location="cpu", weights_only=True)
    raise NotImplementedError(f"reading {path!r}: supply `reader(path, image_size, transform_name=...)` returning [C, T, H, W]")


# frames of pixel-space reference a v2v condition needs: (default, when the clip is >= 64 frames and the condition is "easy")
_V2V_FRAMES = (32, 64)


def collect_references_batch(reference_paths: list, cond_type: str, model_ae: nn.Module, image_size: tuple[int, int],
                             is_causal: bool = False, reader=None) -> list:
    """Per batch item: None (empty path) or the list of VAE-encoded references [C, T', H', W'] its condition uses -
    i2v_head: first frame of the first medium; i2v_tail: last frame of the last; i2v_loop: both; v2v_*: the first / last
    32 (64 for "easy" with enough material; +1 with a causal VAE) frames of the first medium.  Media of one item are
    ';'-separated."""
    reader = reader or _default_reader
    p = next(model_ae.parameters())

    def encode(pixels):
        return model_ae.encode(pixels.unsqueeze(0).to(p.device, p.dtype)).squeeze(0)

    def load(item):
        return reader(item, image_size, transform_name="resize_crop")

    out = []
    for entry in reference_paths:
        if isinstance(entry, str) and entry == "":
            out.append(None)
            continue
        media = entry.split(";") if isinstance(entry, str) else (list(entry) if isinstance(entry, (list, tuple)) else [entry])
        if "v2v" in cond_type:
            clip = load(media[0])
            need = _V2V_FRAMES[1] if (clip.size(1) >= 64 and "easy" in cond_type) else _V2V_FRAMES[0]
            need += int(bool(is_causal))
            assert clip.size(1) >= need, f"need at least {need} reference frames for v2v generation"
            if "head" in cond_type:
                refs = [encode(clip[:, :need])]
            elif "tail" in cond_type:
                refs = [encode(clip[:, -need:])]
            else:
                raise NotImplementedError
        elif cond_type == "i2v_head":
            refs = [encode(load(media[0])[:, :1])]
        elif cond_type == "i2v_tail":
            refs = [encode(load(media[-1])[:, -1:])]
        elif cond_type == "i2v_loop":
            refs = [encode(load(media[0])[:, :1]), encode(load(media[-1])[:, -1:])]
        else:
            raise NotImplementedError(f"Unknown condition type {cond_type}")
        out.append(refs)
    return out


def _conditioned_frames(mask_cond: str, causal: bool):
    """(leading frames, trailing frames, which reference supplies the trailing ones) that a condition pins.  An image-to-video
    tail comes from the item's LAST medium (its last frame), a video-to-video tail from the clip itself (medium 0)."""
    short, long = 8 + int(causal), 16 + int(causal)
    table = {"i2v_head": (1, 0, 0), "i2v_tail": (0, 1, -1), "i2v_loop": (1, 1, -1), "v2v_head": (short, 0, 0),
             "v2v_tail": (0, short, 0), "v2v_head_easy": (long, 0, 0), "v2v_tail_easy": (0, long, 0)}
    if mask_cond not in table:
        assert mask_cond == "t2v", f"Unknown mask condition {mask_cond}"
        return 0, 0, 0
    return table[mask_cond]


def prepare_inference_condition(z: torch.Tensor, mask_cond: str, ref_list: list | None = None, causal: bool = True):
    """inference.py:283-351.  z [B, C, T, H, W] latent noise -> (masks [B, 1, T, H, W], masked_ref [B, C, T, H, W]): ones /
    reference latents on the frames the condition pins, zeros elsewhere.  Leading frames come from the item's first
    reference.  Items without a reference, and single-frame latents, stay unconditioned."""
    B, C, T, H, W = z.shape
    masks = torch.zeros(B, 1, T, H, W)
    masked = torch.zeros(B, C, T, H, W)
    if ref_list is None:
        assert mask_cond == "t2v", f"reference is required for {mask_cond}"
        ref_list = [None] * B
    head, tail, tail_src = _conditioned_frames(mask_cond, causal)
    for i, ref in enumerate(ref_list[:B]):
        if ref is None:
            if mask_cond != "t2v":
                print("no reference found. will default to cond_type t2v!")
            continue
        if T <= 1:
            continue
        if head:
            masks[i, :, :head] = 1
            masked[i, :, :head] = ref[0][:, :head].to(masked.dtype)
        if tail:
            masks[i, :, -tail:] = 1
            masked[i, :, -tail:] = ref[tail_src][:, -tail:].to(masked.dtype)
    return masks.to(z.device, z.dtype), masked.to(z.device, z.dtype)

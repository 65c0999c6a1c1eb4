"""Resolution name + aspect ratio -> (height, width), the arithmetic of `opensora/datasets/aspect.py:4-139` that
`sanitize_sampling_option` / `modify_option_to_t2i` call at inference time: it fixes the latent shape the denoiser and
the VAE see, so it belongs to the data format either side of the hot path.  Bucketing helpers (`bucket_to_shapes`,
`get_closest_ratio`, ...) serve the training dataloader and are not mirrored."""
from __future__ import annotations

import math
import os

# width:height, landscape; the portrait twins are derived (aspect.py:4-14)
ASPECT_RATIO_LD_LIST = ["2.39:1", "2:1", "16:9", "1.85:1", "9:16", "5:8", "3:2", "4:3", "1:1"]


def get_ratio(name: str) -> float:
    w, h = (float(v) for v in name.split(":"))
    return h / w


def _cell() -> int:
    return int(os.environ.get("AE_SPATIAL_COMPRESSION", 16))


def get_aspect_ratios_dict(total_pixels: int = 256 * 256, training: bool = True) -> dict[str, tuple[int, int]]:
    """aspect.py:22-58.  Per listed ratio: the widest width (multiple of the AE cell D) whose area fits, the matching height
    floored to D; in training mode one side may move by one cell when that lands closer to the pixel budget, and sizes
    already produced are skipped.  Every entry also yields its transposed twin ("16:9" -> "9:16" = (width, height)); the
    twins are merged last, so they win over a listed ratio of the same name."""
    D = _cell()
    sizes: dict[str, tuple[int, int]] = {}
    twins: dict[str, tuple[int, int]] = {}
    for name in ASPECT_RATIO_LD_LIST:
        wr, hr = (float(v) for v in name.split(":"))
        width = int(math.sqrt(total_pixels * (wr / hr)) // D) * D
        height = int((total_pixels / width) // D) * D
        if training:
            h0, w0 = height, width
            err = abs(h0 * w0 - total_pixels)
            for h, w in ((h0 - D, w0), (h0 + D, w0), (h0, w0 - D), (h0, w0 + D)):
                if abs(h * w - total_pixels) < err:
                    height, width, err = h, w, abs(h * w - total_pixels)
        if not training or (height, width) not in sizes.values():
            sizes[name] = (height, width)
            twins[":".join(reversed(name.split(":")))] = (width, height)
    sizes.update(twins)
    return sizes


def get_num_pexels(aspect_ratios_dict: dict[str, tuple[int, int]]) -> dict[str, int]:
    return {k: h * w for k, (h, w) in aspect_ratios_dict.items()}


def get_num_tokens(aspect_ratios_dict: dict[str, tuple[int, int]]) -> dict[str, int]:
    D = _cell()
    return {k: h * w // D // D for k, (h, w) in aspect_ratios_dict.items()}


def get_num_pexels_from_name(resolution: str) -> int:
    """"256px" -> 256^2; "720p" -> the 16:9 frame of that height, 720^2 * 16/9 (aspect.py:71-81)."""
    name = resolution.split("_")[0]
    if name.endswith("px"):
        side = int(name[:-2])
        return side * side
    if name.endswith("p"):
        side = int(name[:-1])
        return int(side * side / 9 * 16)
    raise ValueError(f"Invalid resolution {name}")


def get_image_size(resolution: str, ar_ratio: str, training: bool = True) -> tuple[int, int]:
    table = get_aspect_ratios_dict(get_num_pexels_from_name(resolution), training)
    assert ar_ratio in table, f"Aspect ratio {ar_ratio} not found"
    return table[ar_ratio]

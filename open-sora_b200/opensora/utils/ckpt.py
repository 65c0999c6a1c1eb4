"""Checkpoint loading for the drop-in models: the inference-side subset of `opensora/utils/ckpt.py`
(`load_checkpoint` :84-140, `print_load_warning` :64-81, `load_from_hf_hub` :33-47,
`load_from_sharded_state_dict` :50-61) with the same names, argument meaning and key handling, so real Open-Sora
weights (`Open_Sora_v2.safetensors`, `hunyuan_vae.safetensors`) load into the osb200 models unchanged - their
state-dict keys are the reference's.  The training-side half of the reference module (ZeRO master-weight gathering,
async EMA writers, `CheckpointIO`) belongs to the trainer and is out of scope ("out of scope").

Differences, all forced by the environment:
  * no hub download (no network): a path that does not exist locally is looked up in the Hugging Face cache layout
    under `cache_dir` / `$HF_HOME` and raises `FileNotFoundError` otherwise;
  * a sharded checkpoint directory is read through its `*.index.json` weight map (the format ColossalAI's
    `GeneralCheckpointIO` and `transformers` both write) without ColossalAI.
"""
from __future__ import annotations

import glob
import json
import logging
import os

import torch
import torch.nn as nn

_log = logging.getLogger("opensora")


def log_message(*args, level: str = "info") -> None:
    """`opensora/utils/logger.py:72-92`: one line through the package logger."""
    getattr(_log, level if level in ("info", "warning", "error", "debug") else "info")(" ".join(str(a) for a in args))


def load_from_hf_hub(repo_path: str, cache_dir: str | None = None) -> str:
    """`org/repo/file` -> local path of an ALREADY CACHED copy (hub layout
    `<cache>/models--org--repo/snapshots/<rev>/<file>`).  The reference downloads here (ckpt.py:33-47)."""
    parts = repo_path.strip("/").split("/")
    if len(parts) < 3:
        raise FileNotFoundError(f"checkpoint {repo_path!r} does not exist locally and is not an org/repo/file hub path")
    repo, name = "--".join(parts[:2]), "/".join(parts[2:])
    roots = [cache_dir] if cache_dir else []
    home = os.environ.get("HF_HOME", os.path.join(os.path.expanduser("~"), ".cache", "huggingface"))
    roots += [os.environ.get("HF_HUB_CACHE"), os.path.join(home, "hub")]
    for root in filter(None, roots):
        hits = sorted(glob.glob(os.path.join(root, f"models--{repo}", "snapshots", "*", name)))
        if hits:
            return hits[-1]
    raise FileNotFoundError(f"checkpoint {repo_path!r}: not on disk and not in the hub cache ({', '.join(filter(None, roots))}); "
                            "this build has no network access - place the file locally and pass its path")


def print_load_warning(missing: list[str], unexpected: list[str]) -> None:
    """ckpt.py:64-81: report what `load_state_dict(strict=False)` skipped."""
    if missing:
        log_message(f"Got {len(missing)} missing keys:\n\t" + "\n\t".join(missing), level="warning")
    if unexpected:
        log_message(f"Got {len(unexpected)} unexpected keys:\n\t" + "\n\t".join(unexpected), level="warning")
    if not missing and not unexpected:
        log_message("Model loaded successfully")


def _read_file(path: str, device_map="cpu") -> dict:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path, device="cpu")
    sd = torch.load(path, map_location=device_map, weights_only=True)
    return sd


def _rename(sd: dict, rename_keys: dict | None) -> dict:
    """ckpt.py:118-129: the FIRST old prefix found anywhere in a key is replaced (all its occurrences)."""
    if not rename_keys:
        return sd
    out = {}
    for key, v in sd.items():
        new = key
        for old_prefix, new_prefix in rename_keys.items():
            if old_prefix in key:
                new = key.replace(old_prefix, new_prefix)
                break
        out[new] = v
    return out


def load_from_sharded_state_dict(model: nn.Module, ckpt_path: str, model_name: str = "model", strict: bool = False):
    """ckpt.py:50-61: `<ckpt_path>/<model_name>/` holds shard files and an index json whose `weight_map` names the
    shard of every key."""
    root = os.path.join(ckpt_path, model_name)
    if not os.path.isdir(root):
        root = ckpt_path
    index = sorted(glob.glob(os.path.join(root, "*.index.json")))
    if index:
        with open(index[0]) as fh:
            shards = sorted(set(json.load(fh)["weight_map"].values()))
    else:
        shards = sorted(os.path.basename(p) for ext in ("*.safetensors", "*.bin", "*.pt") for p in glob.glob(os.path.join(root, ext)))
    if not shards:
        raise FileNotFoundError(f"no checkpoint shards under {root}")
    sd = {}
    for name in shards:
        sd.update(_read_file(os.path.join(root, name)))
    missing, unexpected = model.load_state_dict(sd, strict=strict)
    print_load_warning(list(missing), list(unexpected))
    return model


def load_checkpoint(model: nn.Module, path: str, cache_dir: str | None = None, device_map: torch.device | str = "cpu",
                    cai_model_name: str = "model", strict: bool = False, rename_keys: dict | None = None) -> nn.Module:
    """ckpt.py:84-140.  Three kinds of checkpoint: a `.safetensors` file (keys optionally renamed), a `.pt` / `.pth`
    torch state dict, or a sharded directory."""
    if not os.path.exists(path):
        log_message(f"Checkpoint not found at {path}, looking in the local Hugging Face cache")
        path = load_from_hf_hub(path, cache_dir)
    log_message(f"Loading checkpoint from {path}")
    if path.endswith(".safetensors"):
        sd = _rename(_read_file(path), rename_keys)
    elif path.endswith((".pt", ".pth", ".bin")):
        sd = _read_file(path, device_map)
    else:
        if not os.path.isdir(path):
            raise ValueError(f"Invalid checkpoint path: {path}")
        return load_from_sharded_state_dict(model, path, model_name=cai_model_name, strict=strict)
    missing, unexpected = model.load_state_dict(sd, strict=strict)
    print_load_warning(list(missing), list(unexpected))
    return model

"""ORACLE tooling — loads the reference's own source files BY PATH from /root/reference (read-only,
present only in the build container, never on the GPU box) so golden vectors can be generated from
the reference's arithmetic itself.  Recipe of SURVEY.md §8(c): empty package shells + 3 stub modules
(`opensora.acceleration.checkpoint.auto_grad_checkpoint`, `opensora.registry.MODELS`,
`opensora.utils.ckpt.load_checkpoint`) and CPU substitutes for the three GPU-only call sites
(flash-attn -> exact SDPA, Liger RMSNorm -> the file's own eager RMSNorm `layers.py:102-111`,
`torch.compile`d timestep_embedding -> its undecorated body).  Nothing here is imported by the
product; only tests/golden/make_*.py use it."""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "opensora"))


def _shell(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = []  # mark as package
    sys.modules[name] = m
    return m


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_mmdit():
    """Returns (layers, math, model) modules of the reference MMDiT, runnable on CPU."""
    import torch
    import torch.nn.functional as F

    saved = {k: v for k, v in sys.modules.items() if k == "opensora" or k.startswith("opensora.")}
    for k in saved:
        del sys.modules[k]
    try:
        for pkg in ("opensora", "opensora.models", "opensora.models.mmdit", "opensora.acceleration", "opensora.utils"):
            _shell(pkg)
        ck = types.ModuleType("opensora.acceleration.checkpoint")
        ck.auto_grad_checkpoint = lambda m, *a, **k: m(*a, **k)
        sys.modules["opensora.acceleration.checkpoint"] = ck
        reg = types.ModuleType("opensora.registry")

        class _R:
            def register_module(self, *a, **k):
                return lambda f: f

        reg.MODELS = _R()
        sys.modules["opensora.registry"] = reg
        uc = types.ModuleType("opensora.utils.ckpt")
        uc.load_checkpoint = lambda model, *a, **k: model
        sys.modules["opensora.utils.ckpt"] = uc

        # torch.compile decorator on timestep_embedding -> identity while the file is executed
        real_compile = torch.compile
        torch.compile = lambda *a, **k: (lambda f: f)
        try:
            base = os.path.join(REF, "opensora", "models", "mmdit")
            math_m = _load("opensora.models.mmdit.math", os.path.join(base, "math.py"))
            layers = _load("opensora.models.mmdit.layers", os.path.join(base, "layers.py"))
            model = _load("opensora.models.mmdit.model", os.path.join(base, "model.py"))
        finally:
            torch.compile = real_compile

        def sdpa_bl_hd(q, k, v):  # flash_attn_func layout: [B, L, H, D]
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
            return o.transpose(1, 2)

        math_m.flash_attn_func = sdpa_bl_hd
        layers.FusedRMSNorm.forward = layers.RMSNorm.forward  # Liger 'llama' mode == eager RMSNorm (layers.py:102-123)
        return layers, math_m, model
    finally:
        for k in [k for k in sys.modules if k == "opensora" or k.startswith("opensora.")]:
            del sys.modules[k]
        sys.modules.update(saved)

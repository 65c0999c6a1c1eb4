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


def load_hunyuan_vae(with_autoencoder: bool = False):
    """Returns (blocks, vae) = the reference's `hunyuan_vae/unet_causal_3d_blocks.py` and `hunyuan_vae/vae.py`
    executed by path, with stand-ins for the third-party `diffusers` pieces they import (absent in this image;
    the reference pins nothing - header says "Modified from diffusers==0.29.2", unet_causal_3d_blocks.py:1):
      * `get_activation("swish"|"silu")` -> nn.SiLU
      * `Attention` -> a restatement of diffusers' AttnProcessor2_0 path with the ctor arguments the reference
        uses (unet_causal_3d_blocks.py:311-325): GroupNorm -> q/k/v Linear -> SDPA(additive mask) -> out Linear
        -> + residual -> / rescale_output_factor.  This piece is therefore NOT pinned by reference source.
    """
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    saved = {k: v for k, v in sys.modules.items()
             if k == "opensora" or k.startswith("opensora.") or k == "diffusers" or k.startswith("diffusers.")}
    for k in saved:
        del sys.modules[k]
    try:
        for pkg in ("opensora", "opensora.models", "opensora.models.hunyuan_vae", "opensora.models.vae",
                    "opensora.acceleration", "diffusers", "diffusers.models", "diffusers.utils"):
            _shell(pkg)
        ck = types.ModuleType("opensora.acceleration.checkpoint")
        ck.auto_grad_checkpoint = lambda m, *a, **k: m(*a, **k)
        ck.checkpoint = lambda fn, *a, use_reentrant=True, **k: fn(*a, **k)
        sys.modules["opensora.acceleration.checkpoint"] = ck

        act = types.ModuleType("diffusers.models.activations")
        act.get_activation = lambda name: nn.SiLU()
        sys.modules["diffusers.models.activations"] = act

        class Attention(nn.Module):
            def __init__(self, query_dim, heads=1, dim_head=64, rescale_output_factor=1.0, eps=1e-6, norm_num_groups=32,
                         spatial_norm_dim=None, residual_connection=True, bias=True, upcast_softmax=True,
                         _from_deprecated_attn_block=True):
                super().__init__()
                inner = heads * dim_head
                self.heads, self.rescale_output_factor, self.residual_connection = heads, rescale_output_factor, residual_connection
                self.group_norm = nn.GroupNorm(num_channels=query_dim, num_groups=norm_num_groups, eps=eps, affine=True)
                self.to_q = nn.Linear(query_dim, inner, bias=bias)
                self.to_k = nn.Linear(query_dim, inner, bias=bias)
                self.to_v = nn.Linear(query_dim, inner, bias=bias)
                self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

            def forward(self, hidden_states, attention_mask=None):
                residual = hidden_states
                B, L, C = hidden_states.shape
                h = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
                q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
                hd = C // self.heads
                q, k, v = (t.view(B, L, self.heads, hd).transpose(1, 2) for t in (q, k, v))
                m = attention_mask
                if m is not None:
                    m = m.view(B, 1, L, L) if m.dim() == 3 else m
                o = F.scaled_dot_product_attention(q, k, v, attn_mask=m, dropout_p=0.0, is_causal=False)
                o = o.transpose(1, 2).reshape(B, L, C)
                o = self.to_out[1](self.to_out[0](o))
                if self.residual_connection:
                    o = o + residual
                return o / self.rescale_output_factor

        ap = types.ModuleType("diffusers.models.attention_processor")
        ap.Attention = Attention
        sys.modules["diffusers.models.attention_processor"] = ap
        du = sys.modules["diffusers.utils"]

        class _Log:
            @staticmethod
            def get_logger(name):
                import logging

                return logging.getLogger(name)

        du.logging = _Log

        class BaseOutput(dict):
            pass

        du.BaseOutput = BaseOutput
        tu = types.ModuleType("diffusers.utils.torch_utils")
        tu.randn_tensor = lambda shape, generator=None, device=None, dtype=None: torch.randn(
            shape, generator=generator, device=device, dtype=dtype)
        sys.modules["diffusers.utils.torch_utils"] = tu
        base = os.path.join(REF, "opensora", "models")
        _load("opensora.models.vae.utils", os.path.join(base, "vae", "utils.py"))
        blocks = _load("opensora.models.hunyuan_vae.unet_causal_3d_blocks",
                       os.path.join(base, "hunyuan_vae", "unet_causal_3d_blocks.py"))
        vae = _load("opensora.models.hunyuan_vae.vae", os.path.join(base, "hunyuan_vae", "vae.py"))
        if not with_autoencoder:
            return blocks, vae
        # autoencoder_kl_causal_3d.py additionally needs the diffusers Mixin plumbing (config registration, hub
        # loading, forward hooks): none of it touches arithmetic, so empty stand-ins suffice.
        _shell("diffusers.loaders")
        _shell("opensora.utils")
        cu = types.ModuleType("diffusers.configuration_utils")
        cu.ConfigMixin = type("ConfigMixin", (), {})
        cu.register_to_config = lambda f: f
        sys.modules["diffusers.configuration_utils"] = cu
        sys.modules["diffusers.loaders"].FromOriginalVAEMixin = type("FromOriginalVAEMixin", (), {})
        for n in ("ADDED_KV_ATTENTION_PROCESSORS", "CROSS_ATTENTION_PROCESSORS", "AttentionProcessor", "AttnAddedKVProcessor",
                  "AttnProcessor"):
            setattr(ap, n, type(n, (), {}))
        mu = types.ModuleType("diffusers.models.modeling_utils")
        mu.ModelMixin = type("ModelMixin", (nn.Module,), {})
        sys.modules["diffusers.models.modeling_utils"] = mu
        au = types.ModuleType("diffusers.utils.accelerate_utils")
        au.apply_forward_hook = lambda f: f
        sys.modules["diffusers.utils.accelerate_utils"] = au
        reg = types.ModuleType("opensora.registry")

        class _R:
            def register_module(self, *a, **k):
                return lambda f: f

        reg.MODELS = _R()
        sys.modules["opensora.registry"] = reg
        uc = types.ModuleType("opensora.utils.ckpt")
        uc.load_checkpoint = lambda model, *a, **k: model
        sys.modules["opensora.utils.ckpt"] = uc
        ae = _load("opensora.models.hunyuan_vae.autoencoder_kl_causal_3d",
                   os.path.join(base, "hunyuan_vae", "autoencoder_kl_causal_3d.py"))
        return blocks, vae, ae
    finally:
        for k in [k for k in sys.modules if k == "opensora" or k.startswith("opensora.") or k == "diffusers"
                  or k.startswith("diffusers.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def load_sampling():
    """The reference's `opensora/utils/sampling.py` executed by path.  Its top-level imports that are absent here
    (mmengine, peft, text encoder, registry, dataset I/O, prompt refinement) are only used by the model-loading / file
    half of the file; they are replaced by empty stand-ins.  Everything the goldens exercise is the reference's own code:
    `time_shift`, `get_schedule`, `pack`, `unpack`, `get_oscillation_gs`, the denoisers, `sanitize_sampling_option` (with
    the real `datasets/aspect.py`), `prepare`, `prepare_ids`, `prepare_api` (with the real `utils/inference.py`:
    `prepare_inference_condition`, `collect_references_batch`).  The two helper modules are returned as attributes
    `ref_aspect` / `ref_inference`."""
    import enum

    names = ("opensora", "opensora.utils", "opensora.datasets", "opensora.models", "opensora.models.mmdit",
             "opensora.models.text", "mmengine", "peft")
    saved = {k: v for k, v in sys.modules.items() if any(k == n or k.startswith(n + ".") for n in ("opensora", "mmengine", "peft"))}
    for k in saved:
        del sys.modules[k]
    try:
        for n in names:
            _shell(n)

        def mod(name, **attrs):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m

        mod("mmengine.config", Config=type("Config", (), {}))
        sys.modules["peft"].PeftModel = type("PeftModel", (), {})
        # `datasets/aspect.py` (math + os only) and `utils/inference.py` are executed for real: the option sanitiser and the
        # conditioning format come from them.  What inference.py imports for file I/O and prompt rewriting is stubbed.
        aspect = _load("opensora.datasets.aspect", os.path.join(REF, "opensora", "datasets", "aspect.py"))
        sys.modules["opensora.datasets"].save_sample = None
        mod("opensora.datasets.utils", read_from_path=None, rescale_image_by_path=None)
        mod("opensora.utils.logger", log_message=print)
        mod("opensora.utils.prompt_refine", refine_prompts=None)
        inference = _load("opensora.utils.inference", os.path.join(REF, "opensora", "utils", "inference.py"))
        mod("opensora.models.mmdit.model", MMDiTModel=type("MMDiTModel", (), {}))
        mod("opensora.models.text.conditioner", HFEmbedder=type("HFEmbedder", (), {}))
        mod("opensora.registry", MODELS=None, build_module=None)
        sampling = _load("opensora.utils.sampling", os.path.join(REF, "opensora", "utils", "sampling.py"))
        sampling.ref_aspect, sampling.ref_inference = aspect, inference
        return sampling
    finally:
        for k in [k for k in sys.modules if any(k == n or k.startswith(n + ".") for n in ("opensora", "mmengine", "peft"))]:
            del sys.modules[k]
        sys.modules.update(saved)

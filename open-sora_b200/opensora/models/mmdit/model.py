"""MMDiTModel / `Flux` factory with the reference's constructor, registry key ("flux"), forward signature and
state-dict keys (`opensora/models/mmdit/model.py:38-303`), running on the osb200 kernels.  Forward-only:
`grad_ckpt_settings` is accepted and ignored (activation checkpointing is training-only)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from opensora.registry import MODELS

from .layers import (DoubleStreamBlock, EmbedND, LastLayer, LigerEmbedND, MLPEmbedder, SingleStreamBlock, _linear,
                     timestep_embedding)


@dataclass
class MMDiTConfig:
    model_type = "MMDiT"
    from_pretrained: str
    cache_dir: str
    in_channels: int
    vec_in_dim: int
    context_in_dim: int
    hidden_size: int
    mlp_ratio: float
    num_heads: int
    depth: int
    depth_single_blocks: int
    axes_dim: list
    theta: int
    qkv_bias: bool
    guidance_embed: bool
    cond_embed: bool = False
    fused_qkv: bool = True
    grad_ckpt_settings: tuple | None = None
    use_liger_rope: bool = False
    patch_size: int = 2

    def get(self, attribute_name, default=None):
        return getattr(self, attribute_name, default)

    def __contains__(self, attribute_name):
        return hasattr(self, attribute_name)


class MMDiTModel(nn.Module):
    config_class = MMDiTConfig

    def __init__(self, config: MMDiTConfig):
        super().__init__()
        self.config = config
        self.in_channels = self.out_channels = config.in_channels
        self.patch_size = config.patch_size
        if config.hidden_size % config.num_heads != 0:
            raise ValueError(f"Hidden size {config.hidden_size} must be divisible by num_heads {config.num_heads}")
        pe_dim = config.hidden_size // config.num_heads
        if sum(config.axes_dim) != pe_dim:
            raise ValueError(f"Got {config.axes_dim} but expected positional dim {pe_dim}")
        self.hidden_size, self.num_heads = config.hidden_size, config.num_heads
        self.pe_embedder = (LigerEmbedND if config.use_liger_rope else EmbedND)(dim=pe_dim, theta=config.theta,
                                                                                 axes_dim=config.axes_dim)
        self.img_in = nn.Linear(self.in_channels, self.hidden_size, bias=True)
        self.time_in = MLPEmbedder(in_dim=256, hidden_dim=self.hidden_size)
        self.vector_in = MLPEmbedder(config.vec_in_dim, self.hidden_size)
        self.guidance_in = MLPEmbedder(in_dim=256, hidden_dim=self.hidden_size) if config.guidance_embed else nn.Identity()
        self.cond_in = nn.Linear(self.in_channels + self.patch_size**2, self.hidden_size, bias=True) \
            if config.cond_embed else nn.Identity()
        self.txt_in = nn.Linear(config.context_in_dim, self.hidden_size)
        self.double_blocks = nn.ModuleList([
            DoubleStreamBlock(self.hidden_size, self.num_heads, mlp_ratio=config.mlp_ratio, qkv_bias=config.qkv_bias,
                              fused_qkv=config.fused_qkv) for _ in range(config.depth)])
        self.single_blocks = nn.ModuleList([
            SingleStreamBlock(self.hidden_size, self.num_heads, mlp_ratio=config.mlp_ratio, fused_qkv=config.fused_qkv)
            for _ in range(config.depth_single_blocks)])
        self.final_layer = LastLayer(self.hidden_size, 1, self.out_channels)
        self.initialize_weights()
        self.forward = self.forward_ckpt  # the reference rebinds forward the same way (model.py:143-146)
        self._input_requires_grad = False
        self._cond_w = None
        self._mod_pack = None
        self._pe_cache = None
        self._sp_group = None
        self.register_load_state_dict_post_hook(lambda m, k: m._drop_caches())

    def _drop_caches(self):
        self._cond_w = self._mod_pack = self._pe_cache = None

    def _apply(self, fn, *a, **k):
        self._drop_caches()
        return super()._apply(fn, *a, **k)

    # ---- per-step constants ------------------------------------------------------------------
    def _grouped_modulation(self, vec: Tensor) -> None:
        """All 2*19 + 38 `Modulation.lin` projections of `vec` as ONE GEMM (the reference launches 76 tiny ones per step,
        layers.py:179-192).  The fp32 result and the column range of every layer ride on `vec`; the processors pick their
        slice (a processor installed on a block this model does not own simply does its own projection)."""
        import osb200

        if self._mod_pack is None:
            lins = [m.lin for b in self.double_blocks for m in (b.img_mod, b.txt_mod)] + [b.modulation.lin for b in self.single_blocks]
            cols, off = {}, 0
            for lin in lins:
                cols[id(lin)] = (off, off + lin.out_features)
                off += lin.out_features
            self._mod_pack = (torch.cat([l.weight for l in lins], 0).contiguous(), torch.cat([l.bias for l in lins], 0).contiguous(), cols)
        w, b, cols = self._mod_pack
        out = osb200.gemm(torch.nn.functional.silu(vec).contiguous(), w, b).float()
        vec._osb_grouped_modulation = (out, cols)

    def _pe(self, txt_ids: Tensor, img_ids: Tensor):
        """`pe_embedder(cat(txt_ids, img_ids))` is step-invariant (utils/sampling.py:437-447 builds the ids once per sample):
        cached on the identity and version of the id tensors the caller passes."""
        key = (id(txt_ids), txt_ids._version, id(img_ids), img_ids._version, tuple(txt_ids.shape), tuple(img_ids.shape), txt_ids.device)
        if self._pe_cache is None or self._pe_cache[0] != key:
            # the tensors are kept alive with the entry, so an id() cannot be recycled while it is the key
            self._pe_cache = (key, self.pe_embedder(torch.cat((txt_ids, img_ids), dim=1)), txt_ids, img_ids)
        return self._pe_cache[1]

    def initialize_weights(self):
        if self.config.cond_embed:
            nn.init.zeros_(self.cond_in.weight)
            nn.init.zeros_(self.cond_in.bias)

    def _lin3(self, x: Tensor, lin: nn.Linear, **kw) -> Tensor:
        """Linear over a [B, L, K] tensor; K is zero-padded to a multiple of 8 when needed (cond_in: K = 68)."""
        B, L, K = x.shape
        x2 = x.to(lin.weight.dtype).reshape(B * L, K)
        w = lin.weight
        if K % 8:
            pad = -K % 8
            x2 = torch.nn.functional.pad(x2, (0, pad))
            if self._cond_w is None or self._cond_w[0] is not lin:
                self._cond_w = (lin, torch.nn.functional.pad(lin.weight, (0, pad)).contiguous())
            w = self._cond_w[1]
        import osb200

        return osb200.gemm(x2.contiguous(), w, lin.bias, **kw).view(B, L, -1)

    def prepare_block_inputs(self, img: Tensor, img_ids: Tensor, txt: Tensor, txt_ids: Tensor, timesteps: Tensor,
                             y_vec: Tensor, cond: Tensor = None, guidance: Tensor | None = None):
        """model.py:154-202."""
        if img.ndim != 3 or txt.ndim != 3:
            raise ValueError("Input img and txt tensors must have 3 dimensions.")
        dt = self.img_in.weight.dtype
        img = self._lin3(img, self.img_in)
        if self.config.cond_embed:
            if cond is None:
                raise ValueError("Didn't get conditional input for conditional model.")
            B, L, C = img.shape
            import osb200

            img = self._lin3(cond, self.cond_in, epilogue=osb200.EPI_BIAS_GATE_RES, residual=img.reshape(B * L, C))
        vec = self.time_in(timestep_embedding(timesteps, 256).to(dt))
        if self.config.guidance_embed:
            if guidance is None:
                raise ValueError("Didn't get guidance strength for guidance distilled model.")
            vec = vec + self.guidance_in(timestep_embedding(guidance, 256).to(dt))
        vec = vec + self.vector_in(y_vec)
        txt = self._lin3(txt, self.txt_in)
        pe = self._pe(txt_ids, img_ids)
        return img, txt, vec, pe

    def enable_sequence_parallel(self, group) -> None:
        """Ulysses sequence parallelism over `group` (the reference's `all_to_all` mode, opensora/models/mmdit/
        distributed.py:473-495,598-634,671-679): the joint txt|img sequence is split into P equal chunks (a rank holds the tail
        of the text and/or a slice of the image tokens), every token-local op runs on the chunk, attention exchanges
        "scatter heads / gather sequence" on q|k|v (ONE all-to-all for the three) and back on the output, and the image
        tokens are all-gathered (var-len) after the final layer.  `None` switches it off."""
        self._sp_group = group

    def _sp_splits(self, Lt: int, Li: int):
        """(P, rank, [txt tokens per rank], [img tokens per rank]) of the equal split of the joint sequence, or None when
        sequence parallelism is off / not applicable (distributed.py:604-617: a rank without image tokens disables it)."""
        import torch.distributed as dist

        g = getattr(self, "_sp_group", None)
        if g is None or not dist.is_initialized() or dist.get_world_size(g) == 1:
            return None
        P, r = dist.get_world_size(g), dist.get_rank(g)
        if (Lt + Li) % P:
            raise ValueError(f"Expected {Lt + Li} % {P} == 0 (distributed.py:600-604)")
        ch = (Lt + Li) // P
        txt_s = [max(0, min((k + 1) * ch, Lt) - min(k * ch, Lt)) for k in range(P)]
        img_s = [ch - t for t in txt_s]
        if 0 in img_s:
            return None
        return P, r, txt_s, img_s

    def forward_ckpt(self, img: Tensor, img_ids: Tensor, txt: Tensor, txt_ids: Tensor, timesteps: Tensor, y_vec: Tensor,
                     cond: Tensor = None, guidance: Tensor | None = None, **kwargs) -> Tensor:
        """model.py:208-233; with a sequence-parallel group: distributed.py:598-681 (forward part)."""
        from . import layers as L_

        img, txt, vec, pe = self.prepare_block_inputs(img, img_ids, txt, txt_ids, timesteps, y_vec, cond, guidance)
        self._grouped_modulation(vec)
        Lt, Li = txt.shape[1], img.shape[1]
        sp = self._sp_splits(Lt, Li)
        vec._osb_txt_len = Lt   # joint positions >= Lt take the image stream's QK-norm weights on every rank
        if sp is not None:
            P, r, txt_s, img_s = sp
            t0, i0 = sum(txt_s[:r]), sum(img_s[:r])
            txt, img = txt[:, t0:t0 + txt_s[r]].contiguous(), img[:, i0:i0 + img_s[r]].contiguous()
        prev = L_._SP["group"]
        L_._SP["group"] = self._sp_group if sp is not None else None
        try:
            for block in self.double_blocks:
                img, txt = block(img, txt, vec, pe)
            x = torch.cat((txt, img), 1)
            for block in self.single_blocks:
                x = block(x, vec, pe)
        finally:
            L_._SP["group"] = prev
        out = self.final_layer(x[:, txt.shape[1]:, ...].contiguous(), vec)
        if sp is not None:
            from opensora.acceleration.communications import gather_forward_split_backward_var_len

            out = gather_forward_split_backward_var_len(out, 1, self._sp_group, sp[3])
        return out

    forward_selective_ckpt = forward_ckpt


@MODELS.register_module("flux")
def Flux(cache_dir: str = None, from_pretrained: str = None, device_map: str | torch.device = "cuda",
         torch_dtype: torch.dtype = torch.bfloat16, strict_load: bool = False, **kwargs) -> MMDiTModel:
    """model.py:271-303.  Weights go through `opensora.utils.ckpt.load_checkpoint` like upstream (safetensors / torch file /
    sharded directory; local or hub-cache paths only - no network here)."""
    config = MMDiTConfig(from_pretrained=from_pretrained, cache_dir=cache_dir, **kwargs)
    model = MMDiTModel(config)
    if from_pretrained:
        from opensora.utils.ckpt import load_checkpoint

        model = load_checkpoint(model, from_pretrained, cache_dir=cache_dir, device_map="cpu", strict=strict_load)
    return model.to(device=device_map, dtype=torch_dtype)

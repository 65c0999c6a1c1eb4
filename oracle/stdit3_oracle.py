"""ORACLE — test infrastructure only, never shipped or measured as the product.

Plain-PyTorch restatement of the STDiT3 denoiser (Open-Sora v1.2 `opensora/models/stdit/stdit3.py`
+ `opensora/models/layers/blocks.py`).  **PARITY UNPINNED**: `/root/reference` (Open-Sora v2.0.0 @
7ad6a96) does not contain this model (`opensora/models/__init__.py:1-5` exports only
dc_ae, hunyuan_vae, mmdit, text, vae); the only in-tree witnesses are `gradio/app.py:119-137` (class
name / module path / call convention) and `docs/report_01.md:11-15`, `docs/report_02.md:20-26`,
`docs/report_03.md:56-82,149-160` (architecture prose).  Every function below therefore cites the
SURVEY.md §8(a-S) / Appendix A row it restates rather than a reference file:line.  The pieces shared
with the in-tree MMDiT (LayerNorm-no-affine + (1+scale)*x+shift, per-head RMSNorm with the cast point
of `mmdit/layers.py:102-111`, GELU-tanh MLP of `layers.py:277-281`, gated residual of
`layers.py:247-252`, exact softmax attention of `mmdit/math.py:22-36`) ARE pinned by reference source
and are cross-checked against it in tests/test_oracle_vs_reference.py via tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class STDiT3Config:
    """SURVEY.md Appendix A 'Registry names / sizes'."""

    input_size: tuple = (None, None, None)
    input_sq_size: int = 512
    in_channels: int = 4
    patch_size: tuple = (1, 2, 2)
    hidden_size: int = 1152
    depth: int = 28
    num_heads: int = 16
    mlp_ratio: float = 4.0
    pred_sigma: bool = True
    caption_channels: int = 4096
    model_max_length: int = 300
    qk_norm: bool = True

    @property
    def out_channels(self) -> int:
        return self.in_channels * 2 if self.pred_sigma else self.in_channels


def STDiT3_XL_2_config(**kw) -> STDiT3Config:
    return STDiT3Config(depth=28, hidden_size=1152, patch_size=(1, 2, 2), num_heads=16, **kw)


def STDiT3_XS_2_config(**kw) -> STDiT3Config:
    """Builder-defined plumbing size (BASELINE.json configs[0]); no upstream equivalent."""
    return STDiT3Config(depth=2, hidden_size=288, patch_size=(1, 2, 2), num_heads=4, **kw)


def t2i_modulate(x, shift, scale):
    """App. A 'Modulation': x * (1 + scale) + shift (same form as mmdit/layers.py:206)."""
    return x * (1 + scale) + shift


def timestep_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """App. A 'Embeddings': sinusoidal cos||sin, fp32 (same construction as mmdit/layers.py:68-88
    without the time_factor)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
            nn.Linear(hidden_size, hidden_size, bias=True),
        )
        self.frequency_embedding_size = frequency_embedding_size

    def forward(self, t: torch.Tensor, dtype) -> torch.Tensor:
        t_freq = timestep_embedding(t, self.frequency_embedding_size).to(dtype)
        return self.mlp(t_freq)


class SizeEmbedder(TimestepEmbedder):
    """fps embedder (App. A: `t = t_embedder(timestep) + fps_embedder(fps)`)."""

    def forward(self, s: torch.Tensor, bs: int) -> torch.Tensor:  # s: [B, 1]
        if s.ndim == 1:
            s = s[:, None]
        if s.shape[0] != bs:
            s = s.repeat(bs // s.shape[0], 1)
        b, dims = s.shape
        s_freq = timestep_embedding(s.reshape(-1), self.frequency_embedding_size).to(self.mlp[0].weight.dtype)
        s_emb = self.mlp(s_freq)
        return s_emb.reshape(b, dims * s_emb.shape[-1])


class Mlp(nn.Module):
    """timm Mlp with GELU(approximate='tanh') (same op as mmdit/layers.py:277-281)."""

    def __init__(self, in_features: int, hidden_features: int, out_features: int | None = None):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU(approximate="tanh")
        self.fc2 = nn.Linear(hidden_features, out_features or in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class CaptionEmbedder(nn.Module):
    """App. A: `y_embedder.y_proj.{fc1,fc2}` + `y_embedding` null caption buffer."""

    def __init__(self, in_channels: int, hidden_size: int, token_num: int):
        super().__init__()
        self.y_proj = Mlp(in_channels, hidden_size, hidden_size)
        self.register_buffer("y_embedding", torch.randn(token_num, in_channels) / in_channels**0.5)

    def forward(self, caption):
        return self.y_proj(caption)


class PositionEmbedding2D(nn.Module):
    """App. A 'Embeddings': 2D sin-cos over (h, w), grid scaled by `scale` and base_size."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim
        half_dim = dim // 2
        inv_freq = 1.0 / (10000 ** (torch.arange(0, half_dim, 2).float() / half_dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)

    def _sin_cos(self, t):
        out = torch.einsum("i,d->id", t, self.inv_freq)
        return torch.cat((torch.sin(out), torch.cos(out)), dim=-1)

    def forward(self, x, h: int, w: int, scale: float = 1.0, base_size: int | None = None):
        dev = self.inv_freq.device
        grid_h = torch.arange(h, device=dev) / scale
        grid_w = torch.arange(w, device=dev) / scale
        if base_size is not None:
            grid_h = grid_h * (base_size / h)
            grid_w = grid_w * (base_size / w)
        grid_h, grid_w = torch.meshgrid(grid_w, grid_h, indexing="ij")  # upstream passes w first
        grid_h = grid_h.t().reshape(-1)
        grid_w = grid_w.t().reshape(-1)
        emb = torch.cat([self._sin_cos(grid_h), self._sin_cos(grid_w)], dim=-1)
        return emb.unsqueeze(0).to(x.dtype)


class PatchEmbed3D(nn.Module):
    """App. A: `x_embedder.proj` Conv3d with kernel = stride = patch (1,2,2)."""

    def __init__(self, patch_size, in_chans: int, embed_dim: int):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        _, _, D, H, W = x.shape
        pt, ph, pw = self.patch_size
        if W % pw:
            x = F.pad(x, (0, pw - W % pw))
        if H % ph:
            x = F.pad(x, (0, 0, 0, ph - H % ph))
        if D % pt:
            x = F.pad(x, (0, 0, 0, 0, 0, pt - D % pt))
        x = self.proj(x)
        return x.flatten(2).transpose(1, 2)  # B, (T H W), C


class LlamaRMSNorm(nn.Module):
    """App. A 'Self-attention': fp32 variance, eps 1e-6, cast back THEN multiply by weight — the cast
    point of the reference's own RMSNorm, mmdit/layers.py:102-111."""

    def __init__(self, hidden_size: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        dt = x.dtype
        x = x.to(torch.float32)
        var = x.pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.variance_epsilon)
        return self.weight * x.to(dt)


class RotaryEmbedding(nn.Module):
    """rotary_embedding_torch.RotaryEmbedding(dim).rotate_queries_or_keys: interleaved-pair rotation,
    theta 1e4, positions 0..L-1 along the sequence axis (App. A 'Self-attention').  Same pairing as the
    reference's `apply_rope` (mmdit/math.py:60-65: pairs (2i, 2i+1))."""

    def __init__(self, dim: int, theta: float = 10000.0):
        super().__init__()
        self.dim = dim
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.register_buffer("freqs", freqs, persistent=False)

    def tables(self, L: int, device=None):
        t = torch.arange(L, device=self.freqs.device if device is None else device, dtype=torch.float32)
        ang = torch.einsum("i,j->ij", t, self.freqs.to(t.device))  # [L, dim/2]
        return ang.cos(), ang.sin()

    def forward(self, x):  # x: [..., L, D]
        L = x.shape[-2]
        cos, sin = self.tables(L, x.device)
        x1, x2 = x[..., 0::2].float(), x[..., 1::2].float()
        out = torch.stack((x1 * cos - x2 * sin, x2 * cos + x1 * sin), dim=-1).flatten(-2)
        return out.to(x.dtype)


class Attention(nn.Module):
    """App. A 'Self-attention' (non-flash branch: softmax in fp32)."""

    def __init__(self, dim: int, num_heads: int, qkv_bias: bool = True, qk_norm: bool = True, rope=None):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim**-0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = LlamaRMSNorm(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = LlamaRMSNorm(self.head_dim) if qk_norm else nn.Identity()
        self.proj = nn.Linear(dim, dim)
        self.rotary_emb = rope

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).view(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q, k = self.q_norm(q), self.k_norm(k)
        if self.rotary_emb is not None:
            q, k = self.rotary_emb(q), self.rotary_emb(k)
        dtype = q.dtype
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = attn.to(torch.float32).softmax(dim=-1).to(dtype)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class MultiHeadCrossAttention(nn.Module):
    """App. A 'Cross-attention': q_linear / kv_linear / proj, block-diagonal mask == per-sample
    key-padding mask; no gate, no QK-norm."""

    def __init__(self, d_model: int, num_heads: int):
        super().__init__()
        self.d_model, self.num_heads, self.head_dim = d_model, num_heads, d_model // num_heads
        self.q_linear = nn.Linear(d_model, d_model)
        self.kv_linear = nn.Linear(d_model, d_model * 2)
        self.proj = nn.Linear(d_model, d_model)

    def forward(self, x, cond, y_lens):
        # x: [B, N, C]; cond: [1, sum(y_lens), C] packed text tokens
        B, N, C = x.shape
        q = self.q_linear(x).view(B, N, self.num_heads, self.head_dim)
        kv = self.kv_linear(cond).view(-1, 2, self.num_heads, self.head_dim)
        outs, off = [], 0
        for b in range(B):
            n = int(y_lens[b])
            k, v = kv[off:off + n, 0], kv[off:off + n, 1]  # [n, H, D]
            off += n
            qb = q[b].transpose(0, 1)  # [H, N, D]
            s = (qb @ k.permute(1, 2, 0)) * (self.head_dim**-0.5)
            p = s.float().softmax(dim=-1).to(qb.dtype)
            outs.append((p @ v.transpose(0, 1)).transpose(0, 1).reshape(N, C))
        return self.proj(torch.stack(outs, 0))


class STDiT3Block(nn.Module):
    """SURVEY.md §8(a-S) row `STDiT3Block.forward`."""

    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0, rope=None, qk_norm=True, temporal=False):
        super().__init__()
        self.temporal = temporal
        self.norm1 = nn.LayerNorm(hidden_size, eps=1e-6, elementwise_affine=False)
        self.attn = Attention(hidden_size, num_heads, qkv_bias=True, qk_norm=qk_norm, rope=rope)
        self.cross_attn = MultiHeadCrossAttention(hidden_size, num_heads)
        self.norm2 = nn.LayerNorm(hidden_size, eps=1e-6, elementwise_affine=False)
        self.mlp = Mlp(hidden_size, int(hidden_size * mlp_ratio))
        self.scale_shift_table = nn.Parameter(torch.randn(6, hidden_size) / hidden_size**0.5)

    @staticmethod
    def t_mask_select(x_mask, x, masked_x, T, S):
        B, N, C = x.shape
        x = x.view(B, T, S, C)
        masked_x = masked_x.view(B, T, S, C)
        return torch.where(x_mask[:, :, None, None], x, masked_x).view(B, N, C)

    def forward(self, x, y, t, y_lens, x_mask=None, t0=None, T=None, S=None):
        B, N, C = x.shape
        sh_msa, sc_msa, g_msa, sh_mlp, sc_mlp, g_mlp = (self.scale_shift_table[None] + t.reshape(B, 6, -1)).chunk(6, dim=1)
        if x_mask is not None:
            sh0_msa, sc0_msa, g0_msa, sh0_mlp, sc0_mlp, g0_mlp = (
                self.scale_shift_table[None] + t0.reshape(B, 6, -1)).chunk(6, dim=1)
        x_m = t2i_modulate(self.norm1(x), sh_msa, sc_msa)
        if x_mask is not None:
            x_m = self.t_mask_select(x_mask, x_m, t2i_modulate(self.norm1(x), sh0_msa, sc0_msa), T, S)
        if self.temporal:
            x_m = x_m.view(B, T, S, C).transpose(1, 2).reshape(B * S, T, C)
            x_m = self.attn(x_m)
            x_m = x_m.view(B, S, T, C).transpose(1, 2).reshape(B, N, C)
        else:
            x_m = self.attn(x_m.view(B * T, S, C)).view(B, N, C)
        x_m_s = g_msa * x_m
        if x_mask is not None:
            x_m_s = self.t_mask_select(x_mask, x_m_s, g0_msa * x_m, T, S)
        x = x + x_m_s
        x = x + self.cross_attn(x, y, y_lens)
        x_m = t2i_modulate(self.norm2(x), sh_mlp, sc_mlp)
        if x_mask is not None:
            x_m = self.t_mask_select(x_mask, x_m, t2i_modulate(self.norm2(x), sh0_mlp, sc0_mlp), T, S)
        x_m = self.mlp(x_m)
        x_m_s = g_mlp * x_m
        if x_mask is not None:
            x_m_s = self.t_mask_select(x_mask, x_m_s, g0_mlp * x_m, T, S)
        return x + x_m_s


class T2IFinalLayer(nn.Module):
    """App. A 'Output'."""

    def __init__(self, hidden_size: int, num_patch: int, out_channels: int):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, num_patch * out_channels, bias=True)
        self.scale_shift_table = nn.Parameter(torch.randn(2, hidden_size) / hidden_size**0.5)

    def forward(self, x, t, x_mask=None, t0=None, T=None, S=None):
        shift, scale = (self.scale_shift_table[None] + t[:, None]).chunk(2, dim=1)
        x_m = t2i_modulate(self.norm_final(x), shift, scale)
        if x_mask is not None:
            sh0, sc0 = (self.scale_shift_table[None] + t0[:, None]).chunk(2, dim=1)
            x_m = STDiT3Block.t_mask_select(x_mask, x_m, t2i_modulate(self.norm_final(x), sh0, sc0), T, S)
        return self.linear(x_m)


class STDiT3(nn.Module):
    """SURVEY.md §8(a-S) row `STDiT3.forward`; module tree / state-dict prefixes of Appendix A."""

    def __init__(self, config: STDiT3Config):
        super().__init__()
        c = self.config = config
        self.hidden_size, self.num_heads, self.depth = c.hidden_size, c.num_heads, c.depth
        self.patch_size, self.in_channels, self.out_channels = c.patch_size, c.in_channels, c.out_channels
        self.input_sq_size = c.input_sq_size
        self.pos_embed = PositionEmbedding2D(c.hidden_size)
        self.rope = RotaryEmbedding(dim=c.hidden_size // c.num_heads)
        self.x_embedder = PatchEmbed3D(c.patch_size, c.in_channels, c.hidden_size)
        self.t_embedder = TimestepEmbedder(c.hidden_size)
        self.fps_embedder = SizeEmbedder(c.hidden_size)
        self.t_block = nn.Sequential(nn.SiLU(), nn.Linear(c.hidden_size, 6 * c.hidden_size, bias=True))
        self.y_embedder = CaptionEmbedder(c.caption_channels, c.hidden_size, c.model_max_length)
        self.spatial_blocks = nn.ModuleList(
            [STDiT3Block(c.hidden_size, c.num_heads, c.mlp_ratio, qk_norm=c.qk_norm) for _ in range(c.depth)])
        self.temporal_blocks = nn.ModuleList(
            [STDiT3Block(c.hidden_size, c.num_heads, c.mlp_ratio, qk_norm=c.qk_norm, temporal=True, rope=self.rope)
             for _ in range(c.depth)])
        self.final_layer = T2IFinalLayer(c.hidden_size, math.prod(c.patch_size), c.out_channels)

    def get_dynamic_size(self, x):
        _, _, T, H, W = x.size()
        pt, ph, pw = self.patch_size
        return -(-T // pt), -(-H // ph), -(-W // pw)

    def encode_text(self, y, mask=None):
        y = self.y_embedder(y)  # [B, 1, L, C]
        if mask is not None:
            if mask.shape[0] != y.shape[0]:
                mask = mask.repeat(y.shape[0] // mask.shape[0], 1)
            mask = mask.reshape(mask.shape[0], -1)
            y = y.squeeze(1).masked_select(mask.unsqueeze(-1) != 0).view(1, -1, self.hidden_size)
            y_lens = mask.sum(dim=1).tolist()
        else:
            y_lens = [y.shape[2]] * y.shape[0]
            y = y.squeeze(1).reshape(1, -1, self.hidden_size)
        return y, y_lens

    def forward(self, x, timestep, y, mask=None, x_mask=None, fps=None, height=None, width=None, **kwargs):
        dtype = self.x_embedder.proj.weight.dtype
        B = x.size(0)
        x, timestep, y = x.to(dtype), timestep.to(dtype), y.to(dtype)
        _, _, Tx, Hx, Wx = x.size()
        T, H, W = self.get_dynamic_size(x)
        S = H * W
        base_size = round(S**0.5)
        resolution_sq = (float(height[0]) * float(width[0])) ** 0.5
        scale = resolution_sq / self.input_sq_size
        pos_emb = self.pos_embed(x, H, W, scale=scale, base_size=base_size)

        t = self.t_embedder(timestep, dtype=x.dtype)
        fps_emb = self.fps_embedder(fps.unsqueeze(1), B)
        t = t + fps_emb
        t_mlp = self.t_block(t)
        t0 = t0_mlp = None
        if x_mask is not None:
            t0 = self.t_embedder(torch.zeros_like(timestep), dtype=x.dtype) + fps_emb
            t0_mlp = self.t_block(t0)

        y, y_lens = self.encode_text(y, mask)

        x = self.x_embedder(x)
        x = (x.view(B, T, S, -1) + pos_emb).view(B, T * S, -1)
        for sb, tb in zip(self.spatial_blocks, self.temporal_blocks):
            x = sb(x, y, t_mlp, y_lens, x_mask, t0_mlp, T, S)
            x = tb(x, y, t_mlp, y_lens, x_mask, t0_mlp, T, S)
        x = self.final_layer(x, t, x_mask, t0, T, S)
        x = self.unpatchify(x, T, H, W, Tx, Hx, Wx)
        return x.to(torch.float32)

    def unpatchify(self, x, N_t, N_h, N_w, R_t, R_h, R_w):
        pt, ph, pw = self.patch_size
        B = x.shape[0]
        x = x.view(B, N_t, N_h, N_w, pt, ph, pw, self.out_channels)
        x = x.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, self.out_channels, N_t * pt, N_h * ph, N_w * pw)
        return x[:, :, :R_t, :R_h, :R_w]


def init_synthetic_weights(model: nn.Module, seed: int = 1234) -> None:
    """Seeded synthetic weights: default init, then overwrite the zero-inits upstream
    uses (temporal attn/mlp out-proj, final layer, cross-attn proj) with N(0, 0.02^2) so every path is
    exercised, and give QK-norm weights / biases non-trivial values."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and "scale_shift_table" not in name:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (0.7 / math.sqrt(fan_in)))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(p.shape[-1]))


def synthetic_inputs(cfg: STDiT3Config, B: int, T: int, H: int, W: int, seed: int = 4321, lens=None):
    """Seeded latents / timesteps / text embeddings of SURVEY.md §8(d)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cfg.in_channels, T, H, W, generator=g)
    timestep = torch.tensor([500.0 + 37.0 * b for b in range(B)])
    y = torch.randn(B, 1, cfg.model_max_length, cfg.caption_channels, generator=g)
    if lens is None:
        lens = [cfg.model_max_length - 11 * b for b in range(B)]
    mask = torch.zeros(B, cfg.model_max_length, dtype=torch.int64)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    fps = torch.full((B,), 24.0)
    height = torch.full((B,), float(H * 8))
    width = torch.full((B,), float(W * 8))
    return dict(x=x, timestep=timestep, y=y, mask=mask, fps=fps, height=height, width=width)

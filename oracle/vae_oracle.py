"""ORACLE — test infrastructure only.  Functional fp32 restatement of the reference's causal 3D VAE
(`opensora/models/hunyuan_vae/{unet_causal_3d_blocks,vae}.py` @ 7ad6a96) over plain NCDHW tensors and a
state-dict-style weight mapping.  PINNED: tests/test_oracle_cpu.py checks it against
tests/golden/vae_blocks.npz, produced by EXECUTING the reference's own source
(tests/golden/make_golden_vae.py via oracle/ref_loader.py).  The mid-block `Attention` is third-party
(diffusers, absent here): restated from its AttnProcessor2_0 arithmetic and therefore unpinned.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def causal_conv3d(x, w, b=None, stride=(1, 1, 1)):
    """unet_causal_3d_blocks.py:63-96 `CausalConv3d.forward`: replicate pad (W k//2,k//2; H k//2,k//2;
    T k-1 front, 0 back) then conv3d without padding.  k=1 -> no padding."""
    k = w.shape[-1]
    if k > 1:
        x = F.pad(x, (k // 2, k // 2, k // 2, k // 2, k - 1, 0), mode="replicate")
    return F.conv3d(x, w, b, stride=stride)


def group_norm_silu(x, w, b, groups=32, eps=1e-6, silu=True):
    """unet_causal_3d_blocks.py:246-250: GroupNorm(32, C, eps 1e-6, affine) then SiLU."""
    y = F.group_norm(x, groups, w, b, eps)
    return F.silu(y) if silu else y


def upsample_causal3d(x, factor=(2, 2, 2)):
    """unet_causal_3d_blocks.py:136-150: nearest; first frame spatial-only, other frames T,H,W -> T' = 1 + ft*(T-1)."""
    T = x.shape[2]
    first = F.interpolate(x[:, :, 0], scale_factor=factor[1:], mode="nearest").unsqueeze(2)
    if T > 1:
        other = F.interpolate(x[:, :, 1:], scale_factor=factor, mode="nearest")
        return torch.cat((first, other), dim=2)
    return first


def resnet_block(W, pfx, x, groups=32):
    """unet_causal_3d_blocks.py:240-259 `ResnetBlockCausal3D.forward` (dropout 0, output_scale_factor 1)."""
    h = group_norm_silu(x, W[pfx + "norm1.weight"], W[pfx + "norm1.bias"], groups)
    h = causal_conv3d(h, W[pfx + "conv1.conv.weight"], W[pfx + "conv1.conv.bias"])
    h = group_norm_silu(h, W[pfx + "norm2.weight"], W[pfx + "norm2.bias"], groups)
    h = causal_conv3d(h, W[pfx + "conv2.conv.weight"], W[pfx + "conv2.conv.bias"])
    if pfx + "conv_shortcut.conv.weight" in W:
        x = causal_conv3d(x, W[pfx + "conv_shortcut.conv.weight"], W.get(pfx + "conv_shortcut.conv.bias"))
    return x + h


def causal_mask(T, hw, device, dtype):
    """unet_causal_3d_blocks.py:52-60: token i sees tokens of frames <= frame(i)."""
    f = torch.arange(T * hw, device=device) // hw
    m = torch.zeros(T * hw, T * hw, device=device, dtype=dtype)
    m.masked_fill_(f[None, :] > f[:, None], float("-inf"))
    return m


def mid_attention(W, pfx, x, groups=32):
    """unet_causal_3d_blocks.py:349-352 + diffusers Attention (1 head of C dims, GroupNorm, bias, residual)."""
    B, C, T, H, Wd = x.shape
    tok = x.permute(0, 2, 3, 4, 1).reshape(B, T * H * Wd, C)
    h = F.group_norm(tok.transpose(1, 2), groups, W[pfx + "group_norm.weight"], W[pfx + "group_norm.bias"], 1e-6).transpose(1, 2)
    q = F.linear(h, W[pfx + "to_q.weight"], W[pfx + "to_q.bias"])
    k = F.linear(h, W[pfx + "to_k.weight"], W[pfx + "to_k.bias"])
    v = F.linear(h, W[pfx + "to_v.weight"], W[pfx + "to_v.bias"])
    s = (q @ k.transpose(1, 2)) * (C ** -0.5) + causal_mask(T, H * Wd, x.device, x.dtype)[None]
    o = torch.softmax(s.float(), dim=-1).to(x.dtype) @ v
    o = F.linear(o, W[pfx + "to_out.0.weight"], W[pfx + "to_out.0.bias"]) + tok
    return o.reshape(B, T, H, Wd, C).permute(0, 4, 1, 2, 3)


def mid_block(W, pfx, x, groups=32):
    """unet_causal_3d_blocks.py:341-355."""
    x = resnet_block(W, pfx + "resnets.0.", x, groups)
    x = mid_attention(W, pfx + "attentions.0.", x, groups)
    return resnet_block(W, pfx + "resnets.1.", x, groups)


def _count(W, pfx):
    n = 0
    while any(k.startswith(f"{pfx}{n}.") for k in W):
        n += 1
    return n


def encoder(W, x, groups=32, strides=None):
    """vae.py:128-150 `EncoderCausal3D.forward`; `strides[i]` = downsample stride of block i or None (vae.py:75-88)."""
    x = causal_conv3d(x, W["conv_in.conv.weight"], W["conv_in.conv.bias"])
    for i in range(_count(W, "down_blocks.")):
        for j in range(_count(W, f"down_blocks.{i}.resnets.")):
            x = resnet_block(W, f"down_blocks.{i}.resnets.{j}.", x, groups)
        key = f"down_blocks.{i}.downsamplers.0.conv.conv.weight"
        if key in W:
            x = causal_conv3d(x, W[key], W[key[:-6] + "bias"], stride=strides[i])
    x = mid_block(W, "mid_block.", x, groups)
    x = group_norm_silu(x, W["conv_norm_out.weight"], W["conv_norm_out.bias"], groups)
    return causal_conv3d(x, W["conv_out.conv.weight"], W["conv_out.conv.bias"])


def decoder(W, z, groups=32, factors=None):
    """vae.py:245-277 `DecoderCausal3D.forward`; `factors[i]` = upsample factor of block i or None (vae.py:199-212)."""
    x = causal_conv3d(z, W["conv_in.conv.weight"], W["conv_in.conv.bias"])
    x = mid_block(W, "mid_block.", x, groups)
    for i in range(_count(W, "up_blocks.")):
        for j in range(_count(W, f"up_blocks.{i}.resnets.")):
            x = resnet_block(W, f"up_blocks.{i}.resnets.{j}.", x, groups)
        key = f"up_blocks.{i}.upsamplers.0.conv.conv.weight"
        if key in W:
            x = upsample_causal3d(x, factors[i])
            x = causal_conv3d(x, W[key], W[key[:-6] + "bias"])
    x = group_norm_silu(x, W["conv_norm_out.weight"], W["conv_norm_out.bias"], groups)
    return causal_conv3d(x, W["conv_out.conv.weight"], W["conv_out.conv.bias"])


def stage_plan(n_blocks=4, time_compression_ratio=4, spatial_compression_ratio=8):
    """Down strides / up factors per block, restating vae.py:66-88 and :190-212."""
    import math

    ns, nt = int(math.log2(spatial_compression_ratio)), int(math.log2(time_compression_ratio))
    down, up = [], []
    for i in range(n_blocks):
        final = i == n_blocks - 1
        sp = i < ns
        td = (i >= n_blocks - 1 - nt) and not final
        down.append((2 if td else 1, 2 if sp else 1, 2 if sp else 1) if (sp or td) else None)
        tu = (i >= n_blocks - 1 - nt) and not final
        up.append((2 if tu else 1, 2 if sp else 1, 2 if sp else 1) if (sp or tu) else None)
    return down, up


# ---- tiled / blended modes of AutoencoderKLCausal3D (autoencoder_kl_causal_3d.py:360-552) -------------------------
def _blend(a, b, extent, dim):
    """:360-382 blend_v / blend_h / blend_t (in place on b, slice by slice in the reference)."""
    extent = min(a.shape[dim], b.shape[dim], extent)
    for k in range(extent):
        sb, sa = b.select(dim, k), a.select(dim, a.shape[dim] - extent + k)
        sb.copy_(sa * (1 - k / extent) + sb * (k / extent))
    return b


def _stitch(rows, blend, limit):
    out_rows = []
    for i, row in enumerate(rows):
        out = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = _blend(rows[i - 1][j], tile, blend, -2)
            if j > 0:
                tile = _blend(row[j - 1], tile, blend, -1)
            out.append(tile[:, :, :, :limit, :limit])
        out_rows.append(torch.cat(out, dim=-1))
    return torch.cat(out_rows, dim=-2)


def _stitch_t(row, blend, t_limit):
    out = []
    for i, tile in enumerate(row):
        if i > 0:
            tile = _blend(row[i - 1], tile, blend, 2)
            out.append(tile[:, :, :t_limit])
        else:
            out.append(tile[:, :, : t_limit + 1])
    return torch.cat(out, dim=2)


def tiled_autoencoder(enc_fn, dec_fn, sample_size, sample_tsize, n_blocks=4, time_ratio=4, overlap=0.25, spatial=True, temporal=True):
    """Returns (encode_moments(x), decode(z)) reproducing :269-335 dispatch + :384-552 tiling around per-tile callables
    enc_fn(x_tile) -> moments and dec_fn(z_tile) -> sample."""
    ls = int(sample_size / (2 ** (n_blocks - 1)))
    lt = sample_tsize // time_ratio

    def sp_enc(x):
        ov, bl = int(sample_size * (1 - overlap)), int(ls * overlap)
        rows = [[enc_fn(x[:, :, :, i:i + sample_size, j:j + sample_size]) for j in range(0, x.shape[-1], ov)]
                for i in range(0, x.shape[-2], ov)]
        return _stitch(rows, bl, ls - bl)

    def sp_dec(z):
        ov, bl = int(ls * (1 - overlap)), int(sample_size * overlap)
        rows = [[dec_fn(z[:, :, :, i:i + ls, j:j + ls]) for j in range(0, z.shape[-1], ov)] for i in range(0, z.shape[-2], ov)]
        return _stitch(rows, bl, sample_size - bl)

    def encode(x):
        if temporal and x.shape[2] > sample_tsize:
            ov, bl = int(sample_tsize * (1 - overlap)), int(lt * overlap)
            row = []
            for i in range(0, x.shape[2], ov):
                tile = x[:, :, i:i + sample_tsize + 1]
                big = spatial and (tile.shape[-1] > sample_size or tile.shape[-2] > sample_size)
                tile = sp_enc(tile) if big else enc_fn(tile)
                row.append(tile[:, :, 1:] if i > 0 else tile)
            return _stitch_t(row, bl, lt - bl)
        if spatial and (x.shape[-1] > sample_size or x.shape[-2] > sample_size):
            return sp_enc(x)
        return enc_fn(x)

    def decode(z):
        if temporal and z.shape[2] > lt:
            ov, bl = int(lt * (1 - overlap)), int(sample_tsize * overlap)
            row = []
            for i in range(0, z.shape[2], ov):
                tile = z[:, :, i:i + lt + 1]
                big = spatial and (tile.shape[-1] > ls or tile.shape[-2] > ls)
                dec = sp_dec(tile) if big else dec_fn(tile)
                row.append(dec[:, :, 1:] if i > 0 else dec)
            return _stitch_t(row, bl, sample_tsize - bl)
        if spatial and (z.shape[-1] > ls or z.shape[-2] > ls):
            return sp_dec(z)
        return dec_fn(z)

    return encode, decode

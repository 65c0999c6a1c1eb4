"""EncoderCausal3D / DecoderCausal3D / DiagonalGaussianDistribution with the reference's names, constructor
arguments and state-dict keys (`opensora/models/hunyuan_vae/vae.py:40-150,153-277,280-340`), running on the
osb200 kernels.  Encoder/decoder take and return **NDHWC bf16** (the NCDHW<->NDHWC conversion happens once at
`AutoencoderKLCausal3D.encode/decode`)."""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
import torch.nn as nn

from opensora.acceleration.communications import gather_forward_split_backward_var_len

from .unet_causal_3d_blocks import (CausalConv3d, DownEncoderBlockCausal3D, UNetMidBlockCausal3D, UpDecoderBlockCausal3D,
                                    frame_partition, temporal_shard)


def _plan(n_blocks, time_compression_ratio, spatial_compression_ratio):
    """(down strides, up factors) per block: vae.py:66-88 and :190-212."""
    ns, nt = int(math.log2(spatial_compression_ratio)), int(math.log2(time_compression_ratio))
    if time_compression_ratio not in (4, 8):
        raise ValueError(f"Unsupported time_compression_ratio: {time_compression_ratio}.")
    out = []
    for i in range(n_blocks):
        final = i == n_blocks - 1
        sp = i < ns
        tm = ((i >= n_blocks - 1 - nt) and not final) if time_compression_ratio == 4 else sp
        out.append((2 if tm else 1, 2 if sp else 1, 2 if sp else 1) if (sp or tm) else None)
    return out


class EncoderCausal3D(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(64,), layers_per_block=2, norm_num_groups=32,
                 act_fn="silu", double_z=True, mid_block_add_attention=True, time_compression_ratio=4,
                 spatial_compression_ratio=8, dropout=0.0):
        super().__init__()
        self.layers_per_block = layers_per_block
        self.conv_in = CausalConv3d(in_channels, block_out_channels[0], kernel_size=3, stride=1)
        self.down_blocks = nn.ModuleList([])
        plan = _plan(len(block_out_channels), time_compression_ratio, spatial_compression_ratio)
        oc = block_out_channels[0]
        for i, ch in enumerate(block_out_channels):
            ic, oc = oc, ch
            self.down_blocks.append(DownEncoderBlockCausal3D(
                num_layers=layers_per_block, in_channels=ic, out_channels=oc, add_downsample=plan[i] is not None,
                downsample_stride=plan[i] or 1, resnet_eps=1e-6, resnet_groups=norm_num_groups))
        self.mid_block = UNetMidBlockCausal3D(in_channels=block_out_channels[-1], resnet_eps=1e-6, output_scale_factor=1,
                                              attention_head_dim=block_out_channels[-1], resnet_groups=norm_num_groups,
                                              add_attention=mid_block_add_attention)
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[-1], num_groups=norm_num_groups, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = CausalConv3d(block_out_channels[-1], 2 * out_channels if double_z else out_channels, kernel_size=3)

        self.time_downsample = time_compression_ratio
        self.shard_group = None   # set by AutoencoderKLCausal3D.enable_temporal_parallel

    def forward(self, sample):  # NDHWC bf16 (channels padded to a multiple of 8)
        group = self.shard_group
        f = self.time_downsample
        latent_frames = (sample.shape[1] - 1) // f + 1
        if group is None or (sample.shape[1] - 1) % f or latent_frames < 2 * dist.get_world_size(group):
            x = self.conv_in(sample)
            for blk in self.down_blocks:
                x = blk(x)
            x = self.mid_block(x)
            return self.conv_out(x, norm=self.conv_norm_out, silu=True)
        # Frame-sharded encode (the mirror image of DecoderCausal3D's): rank r takes the pixel frames of its run of LATENT
        # frames (rank 0: f c_0 - (f - 1), the others f c_r), so every temporally strided stage ends on a shard boundary.
        # Every layer runs on the local frames: convolutions with the causal halo from the left neighbour and group-wide
        # GroupNorm statistics, the mid block's frame-causal attention with local queries against the gathered keys / values;
        # one gather of the latent at the end.
        P, r = dist.get_world_size(group), dist.get_rank(group)
        counts = frame_partition(latent_frames, P)
        pixel = [f * c - (f - 1 if q == 0 else 0) for q, c in enumerate(counts)]
        first = sum(pixel[:r])
        x = sample[:, first:first + pixel[r]].contiguous()
        with temporal_shard(group, start=first):
            x = self.conv_in(x)
            for blk in self.down_blocks:
                x = blk(x)
            assert x.shape[1] == counts[r], (x.shape, counts, r)
            x = self.mid_block(x)
            x = self.conv_out(x, norm=self.conv_norm_out, silu=True)
        return gather_forward_split_backward_var_len(x, 1, group, counts)


class DecoderCausal3D(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(64,), layers_per_block=2, norm_num_groups=32,
                 act_fn="silu", mid_block_add_attention=True, time_compression_ratio=4, spatial_compression_ratio=8,
                 dropout=0.0):
        super().__init__()
        self.layers_per_block = layers_per_block
        self.conv_in = CausalConv3d(in_channels, block_out_channels[-1], kernel_size=3, stride=1)
        self.mid_block = UNetMidBlockCausal3D(in_channels=block_out_channels[-1], resnet_eps=1e-6, output_scale_factor=1,
                                              attention_head_dim=block_out_channels[-1], resnet_groups=norm_num_groups,
                                              add_attention=mid_block_add_attention)
        self.up_blocks = nn.ModuleList([])
        rev = list(reversed(block_out_channels))
        plan = _plan(len(block_out_channels), time_compression_ratio, spatial_compression_ratio)
        oc = rev[0]
        for i, ch in enumerate(rev):
            pc, oc = oc, ch
            self.up_blocks.append(UpDecoderBlockCausal3D(
                num_layers=layers_per_block + 1, in_channels=pc, out_channels=oc, add_upsample=plan[i] is not None,
                upsample_scale_factor=plan[i] or (1, 1, 1), resnet_eps=1e-6, resnet_groups=norm_num_groups))
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=norm_num_groups, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = CausalConv3d(block_out_channels[0], out_channels, kernel_size=3)

        self.time_upsample = time_compression_ratio
        self.shard_group = None   # set by AutoencoderKLCausal3D.enable_temporal_parallel

    def forward(self, sample):  # NDHWC bf16 latent
        group = self.shard_group
        # fewer than 2 latent frames per rank (the causal halo is 2 frames deep): every rank decodes the whole (tile of the)
        # latent - same result on every rank, no exchange
        if group is None or sample.shape[1] < 2 * dist.get_world_size(group):
            x = self.conv_in(sample)
            x = self.mid_block(x)
            for blk in self.up_blocks:
                x = blk(x)
            return self.conv_out(x, norm=self.conv_norm_out, silu=True)
        # Frame-sharded decode: each rank keeps its run of latent frames through EVERY layer - conv_in, the mid block (frame-
        # causal attention: local queries against the gathered keys / values of the earlier frames), the up blocks, where
        # the activations grow 4 x 8 x 8 fold, and conv_out - and the pixel frames are gathered once at the end.
        P, r = dist.get_world_size(group), dist.get_rank(group)
        counts = frame_partition(sample.shape[1], P)
        first = sum(counts[:r])
        x = sample[:, first:first + counts[r]].contiguous()
        with temporal_shard(group, start=first):
            x = self.conv_in(x)
            x = self.mid_block(x)
            for blk in self.up_blocks:
                x = blk(x)
            x = self.conv_out(x, norm=self.conv_norm_out, silu=True)
        f = self.time_upsample
        out_frames = [f * c - (f - 1 if q == 0 else 0) for q, c in enumerate(counts)]   # the first frame is not repeated
        assert x.shape[1] == out_frames[r], (x.shape, out_frames, r)
        return gather_forward_split_backward_var_len(x, 1, group, out_frames)


class DiagonalGaussianDistribution:
    """vae.py:280-340.  The posterior over the latent: `parameters` holds mean | logvar along the channel axis (axis 1 for
    image / video tensors, the last axis for token tensors [B, L, C]).  Elementwise work on a latent-sized tensor, kept in
    torch."""

    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        if parameters.ndim == 3:
            axis = 2
        elif parameters.ndim in (4, 5):
            axis = 1
        else:
            raise NotImplementedError(f"posterior parameters with {parameters.ndim} dimensions")
        self.parameters = parameters
        self.deterministic = deterministic
        self.mean, logvar = parameters.chunk(2, dim=axis)
        self.logvar = logvar.clamp(-30.0, 20.0)
        if deterministic:
            self.std = self.var = torch.zeros_like(self.mean)
        else:
            self.std, self.var = (0.5 * self.logvar).exp(), self.logvar.exp()

    def sample(self, generator=None) -> torch.Tensor:
        eps = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * eps

    def kl(self, other=None) -> torch.Tensor:
        if self.deterministic:
            return torch.Tensor([0.0])
        axes = list(range(1, self.mean.ndim))
        if other is None:   # against N(0, I)
            terms = self.mean.pow(2) + self.var - 1.0 - self.logvar
        else:
            terms = (self.mean - other.mean).pow(2) / other.var + self.var / other.var - 1.0 - self.logvar + other.logvar
        return 0.5 * terms.sum(dim=axes)

    def nll(self, sample: torch.Tensor, dims=(1, 2, 3)) -> torch.Tensor:
        if self.deterministic:
            return torch.Tensor([0.0])
        import math

        return 0.5 * (math.log(2.0 * math.pi) + self.logvar + (sample - self.mean).pow(2) / self.var).sum(dim=list(dims))

    def mode(self) -> torch.Tensor:
        return self.mean

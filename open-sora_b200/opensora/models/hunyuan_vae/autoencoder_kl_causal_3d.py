"""`AutoencoderKLCausal3D` / `CausalVAE3D_HUNYUAN` with the reference's public surface
(`opensora/models/hunyuan_vae/autoencoder_kl_causal_3d.py:59-146,269-358,554-638`): registry key
`"hunyuan_vae"`, `encode / decode / forward / get_latent_size`, `scale_factor`, `shift_factor`, `z_channels`,
compression ratios, tiling toggles, identical state-dict keys — arithmetic on libosb200 (sm_100a).

Round-1 scope: the UNTILED path (`:298-304`, `:318-335`).  B200's 180 GB holds a 65x720x1280 decode untiled, which is
also the faster path (no ~1.65x tile-overlap recompute); the tiled/blended mode (`:384-552`), whose numerics differ from
untiled by construction (per-tile GroupNorm statistics), is SURVEY.md §8(f)-3 and raises NotImplementedError when a
caller enables it AND the input exceeds the tile thresholds."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from opensora.registry import MODELS

from .vae import DecoderCausal3D, DiagonalGaussianDistribution, EncoderCausal3D


@dataclass
class AutoEncoder3DConfig:
    from_pretrained: str | None = None
    act_fn: str = "silu"
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 16
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scale_factor: float = 0.476986
    shift_factor: float = 0
    time_compression_ratio: int = 4
    spatial_compression_ratio: int = 8
    mid_block_add_attention: bool = True
    block_out_channels: tuple = (128, 256, 512, 512)
    sample_size: int = 256
    sample_tsize: int = 64
    use_slicing: bool = False
    use_spatial_tiling: bool = False
    use_temporal_tiling: bool = False
    tile_overlap_factor: float = 0.25
    dropout: float = 0.0
    channel: bool = False


class AutoencoderKLCausal3D(nn.Module):
    def __init__(self, config: AutoEncoder3DConfig):
        super().__init__()
        self.config = config
        self.scale_factor, self.shift_factor = config.scale_factor, config.shift_factor
        self.time_compression_ratio = config.time_compression_ratio
        self.spatial_compression_ratio = config.spatial_compression_ratio
        self.z_channels = config.latent_channels
        common = dict(block_out_channels=config.block_out_channels, layers_per_block=config.layers_per_block,
                      act_fn=config.act_fn, norm_num_groups=config.norm_num_groups,
                      time_compression_ratio=config.time_compression_ratio,
                      spatial_compression_ratio=config.spatial_compression_ratio,
                      mid_block_add_attention=config.mid_block_add_attention, dropout=config.dropout)
        self.encoder = EncoderCausal3D(in_channels=config.in_channels, out_channels=config.latent_channels, double_z=True, **common)
        self.decoder = DecoderCausal3D(in_channels=config.latent_channels, out_channels=config.out_channels, **common)
        self.quant_conv = nn.Conv3d(2 * config.latent_channels, 2 * config.latent_channels, kernel_size=1)
        self.post_quant_conv = nn.Conv3d(config.latent_channels, config.latent_channels, kernel_size=1)
        self.use_slicing = config.use_slicing
        self.use_spatial_tiling = config.use_spatial_tiling
        self.use_temporal_tiling = config.use_temporal_tiling
        self.tile_sample_min_tsize = config.sample_tsize
        self.tile_latent_min_tsize = config.sample_tsize // config.time_compression_ratio
        ss = config.sample_size[0] if isinstance(config.sample_size, (list, tuple)) else config.sample_size
        self.tile_sample_min_size = config.sample_size
        self.tile_latent_min_size = int(ss / (2 ** (len(config.block_out_channels) - 1)))
        self.tile_overlap_factor = config.tile_overlap_factor

    # ---- toggles (same names as the reference :148-189) ------------------------------------------------
    def enable_temporal_tiling(self, use_tiling: bool = True):
        self.use_temporal_tiling = use_tiling

    def disable_temporal_tiling(self):
        self.enable_temporal_tiling(False)

    def enable_spatial_tiling(self, use_tiling: bool = True):
        self.use_spatial_tiling = use_tiling

    def disable_spatial_tiling(self):
        self.enable_spatial_tiling(False)

    def enable_tiling(self, use_tiling: bool = True):
        self.enable_spatial_tiling(use_tiling)
        self.enable_temporal_tiling(use_tiling)

    def disable_tiling(self):
        self.enable_tiling(False)

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    # ---- helpers -------------------------------------------------------------------------------------
    def _check(self):
        import osb200

        w = self.quant_conv.weight
        if not w.is_cuda or w.dtype != torch.bfloat16:
            raise osb200.OsbError("AutoencoderKLCausal3D (osb200) runs on CUDA in bfloat16 only; no CPU / eager fallback")
        return osb200

    def _tiled(self, t, h, w, latent: bool):
        mt = self.tile_latent_min_tsize if latent else self.tile_sample_min_tsize
        ms = self.tile_latent_min_size if latent else self.tile_sample_min_size
        if (self.use_temporal_tiling and t > mt) or (self.use_spatial_tiling and (h > ms or w > ms)):
            raise NotImplementedError("tiled VAE encode/decode (autoencoder_kl_causal_3d.py:384-552) is SURVEY.md §8(f)-3; "
                                      "call disable_tiling(): the untiled path fits in B200 HBM")

    @staticmethod
    def _to_ndhwc(x, cpad=None):
        x = x.permute(0, 2, 3, 4, 1)
        if cpad is not None and x.shape[-1] < cpad:
            x = torch.nn.functional.pad(x, (0, cpad - x.shape[-1]))
        return x.contiguous()

    @staticmethod
    def _to_ncdhw(x):
        return x.permute(0, 4, 1, 2, 3).contiguous()

    # ---- public API (:269-358, :554-622) -------------------------------------------------------------------
    def encode(self, x, sample_posterior=True, return_posterior=False, generator=None):
        osb = self._check()
        assert x.dim() == 5, "The input tensor should have 5 dimensions."
        self._tiled(x.shape[2], x.shape[3], x.shape[4], latent=False)
        dev = self.quant_conv.weight.device
        xin = self._to_ndhwc(x.to(dev, torch.bfloat16), cpad=8)
        h = self.encoder(xin)                                                  # [B,T',h,w,2*latent]
        nb, T, H, W, C2 = h.shape
        qw = self.quant_conv.weight.reshape(C2, C2)
        moments = osb.gemm(h.reshape(-1, C2), qw, self.quant_conv.bias).view(nb, T, H, W, C2)
        posterior = DiagonalGaussianDistribution(self._to_ncdhw(moments))
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        z = self.scale_factor * (z - self.shift_factor)
        return (z, posterior) if return_posterior else z

    def decode(self, z):
        osb = self._check()
        self._tiled(z.shape[2], z.shape[3], z.shape[4], latent=True)
        dev = self.quant_conv.weight.device
        z = z.to(dev, torch.bfloat16) / self.scale_factor + self.shift_factor
        zl = self._to_ndhwc(z)
        nb, T, H, W, C = zl.shape
        pw = self.post_quant_conv.weight.reshape(C, C)
        zl = osb.gemm(zl.reshape(-1, C), pw, self.post_quant_conv.bias).view(nb, T, H, W, C)
        return self._to_ncdhw(self.decoder(zl))

    def forward(self, x, sample_posterior=True, generator=None):
        z, posterior = self.encode(x, return_posterior=True, sample_posterior=sample_posterior, generator=generator)
        return self.decode(z), posterior, z

    def get_last_layer(self):
        return self.decoder.conv_out.conv.weight

    def get_latent_size(self, input_size: list[int]) -> list[int]:
        """autoencoder_kl_causal_3d.py:615-622."""
        latent_size = [(input_size[0] - 1) // self.time_compression_ratio + 1]
        for i in range(1, 3):
            latent_size.append((input_size[i] - 1) // self.spatial_compression_ratio + 1)
        return latent_size


@MODELS.register_module("hunyuan_vae")
def CausalVAE3D_HUNYUAN(from_pretrained: str = None, device_map: str | torch.device = "cuda",
                        torch_dtype: torch.dtype = torch.bfloat16, **kwargs) -> AutoencoderKLCausal3D:
    """Same factory signature as autoencoder_kl_causal_3d.py:625-638."""
    fields = AutoEncoder3DConfig.__dataclass_fields__
    config = AutoEncoder3DConfig(from_pretrained=from_pretrained, **{k: v for k, v in kwargs.items() if k in fields})
    model = AutoencoderKLCausal3D(config)
    if from_pretrained:
        if from_pretrained.endswith(".safetensors"):
            from safetensors.torch import load_file

            sd = load_file(from_pretrained)
        else:
            sd = torch.load(from_pretrained, map_location="cpu")
        model.load_state_dict(sd, strict=True)
    return model.to(device=device_map, dtype=torch_dtype) if torch.cuda.is_available() or str(device_map) == "cpu" \
        else model.to(dtype=torch_dtype)

from .autoencoder_kl_causal_3d import AutoEncoder3DConfig, AutoencoderKLCausal3D, CausalVAE3D_HUNYUAN  # noqa: F401

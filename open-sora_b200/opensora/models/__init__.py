from . import hunyuan_vae, mmdit, stdit  # noqa: F401  (registers "hunyuan_vae", "STDiT3-XL/2", ... in opensora.registry.MODELS)

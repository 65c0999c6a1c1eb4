from . import stdit  # noqa: F401  (registers STDiT3-XL/2 etc. in opensora.registry.MODELS)

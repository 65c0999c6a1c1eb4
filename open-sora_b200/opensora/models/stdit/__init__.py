from .stdit3 import STDiT3, STDiT3Config, STDiT3_3B_2, STDiT3_XL_2, STDiT3_XS_2  # noqa: F401

from .layers import (DoubleStreamBlock, DoubleStreamBlockProcessor, EmbedND, LastLayer, LigerEmbedND, MLPEmbedder,  # noqa: F401
                     SingleStreamBlock, SingleStreamBlockProcessor, timestep_embedding)
from .model import Flux, MMDiTConfig, MMDiTModel  # noqa: F401

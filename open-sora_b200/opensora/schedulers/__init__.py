from .rf import RFLOW, timestep_transform  # noqa: F401

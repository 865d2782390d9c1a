"""interactive_unet_b200: B200-native volume prediction behind interactive-unet's Python entry points.

The directory is named `interactive-unet_b200/` (repo layout contract); it is imported as
`interactive_unet_b200` through the shim package of that name at the repo root.

    from interactive_unet_b200 import predict, unet      # same names as the reference's modules
"""
from . import _lib, distributed, engine, network, predict, unet  # noqa: F401
from .engine import Engine, gaussian_window_1d  # noqa: F401
from .unet import UNet  # noqa: F401

__all__ = ["Engine", "UNet", "predict", "unet", "engine", "network", "gaussian_window_1d"]

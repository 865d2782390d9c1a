"""Import the reference's own `predict.py` unmodified.  TEST INFRASTRUCTURE.

`/root/reference/interactive_unet/predict.py` imports `zarr`, and through
`utils.py` / `unet.py` also `tifffile`, `skimage`, `lightning` and
`segmentation_models_pytorch`; none of them is installed in this image
(SURVEY.md App. C).  The functions on the hot path never touch those
modules (`predict_block` `predict.py:79-112`, `gaussian_3d` `:327-347`,
`get_block_coordinates` `:362-411`, `get_padded_block` `:291-316`,
`get_shard_coordinates` `:318-325`, `find_max_batch_size` `:49-77`), so
placeholder modules are enough to let the file import and the functions run
exactly as written.

The reference tree only exists in the build container; `available()` is False
on the GPU box and callers must then use `oracle.predict_port` + the committed
fixtures instead.
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("IU_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "interactive_unet", "predict.py"))


def _placeholder(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__placeholder__ = True
    return mod


class _LightningModuleStandIn(torch.nn.Module):
    """What `unet.py:9,23` and `predict.py:53,83` use of LightningModule."""

    def save_hyperparameters(self, *a, **k):
        return None

    def log(self, *a, **k):
        return None

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


def _install_placeholders():
    def _unavailable(*a, **k):
        raise RuntimeError("placeholder module: not installed in this image")

    wanted = {
        "zarr": dict(open=_unavailable),
        "tifffile": dict(imread=_unavailable, imwrite=_unavailable),
        "skimage": dict(),
        "skimage.io": dict(imsave=_unavailable, imread=_unavailable),
        "lightning": dict(LightningModule=_LightningModuleStandIn),
        "segmentation_models_pytorch": dict(),
    }
    for name, attrs in wanted.items():
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _placeholder(name, **attrs)
    if getattr(sys.modules["skimage"], "__placeholder__", False):
        sys.modules["skimage"].io = sys.modules["skimage.io"]


_cached = None


def load():
    """Return the reference's `interactive_unet.predict` module, imported verbatim."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not present under {REFERENCE_ROOT}")
    _install_placeholders()
    # the product's checkpoint-unpickling shim may have registered a path-less stand-in package
    if not hasattr(sys.modules.get("interactive_unet"), "__path__"):
        for name in [m for m in sys.modules if m == "interactive_unet" or m.startswith("interactive_unet.")]:
            del sys.modules[name]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _cached = importlib.import_module("interactive_unet.predict")
    return _cached

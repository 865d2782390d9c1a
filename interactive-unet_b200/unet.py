"""Drop-in for the reference's `UNet` module (`/root/reference/interactive_unet/unet.py:10-69`).

Same constructor arguments, same `forward` contract (fp32 `[B,1,H,W]` in `[0,1]` -> softmax
probabilities fp32 `[B,C,H,W]`), same `state_dict()` keys (`model.<smp key>`), `.device`, `.eval()`,
`.to()`, `load_from_checkpoint(checkpoint_path=...)`, `configure_optimizers`, `training_step`,
`validation_step` and `_log_metrics` (`unet.py:71-116`).  In eval mode `forward` runs entirely in the
native sm_100a engine; there is no PyTorch or CPU fallback for inference -- a missing library or a
non-CUDA input raises.  Training mode runs the same parameters through stock autograd, so the
reference's trainer (`trainer.py:37-49`) can fit this module; validation steps (eval mode) use the engine.

Accelerated configurations: `architecture='U-Net'` with `encoder_name` `'resnet34'` or `'resnet18'`,
`num_channels=1`.  DELIBERATE DIFFERENCE from `unet.py:15-20`: the
defaults here are `encoder_name='resnet34', pretrained=False` (the reference's are `'mit_b0', True`: a
transformer encoder and a weight download, neither of which this engine has).  Every other
configuration raises `NotImplementedError` naming the stock module to use instead -- it is never run on
another code path silently.
"""
import sys
import types

import torch
import torch.nn as nn

from .engine import Engine
from .network import ENCODER_BLOCKS, SmpUnetResnet34

try:                                    # the reference subclasses LightningModule (unet.py:9)
    import lightning as _L
    _Base = _L.LightningModule
except Exception:                       # lightning is not installed in this image
    _L = None

    class _Base(nn.Module):
        def save_hyperparameters(self, *args, **kwargs):
            return None

        def log(self, *args, **kwargs):
            return None

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")


def _install_pickle_shims():
    """Reference checkpoints pickle `loss_function=interactive_unet.metrics.<fn>` (unet.py:17,23).
    Make that module path resolvable so `torch.load` can unpickle hyper-parameters."""
    if "interactive_unet.metrics" in sys.modules:
        return
    try:
        import interactive_unet.metrics  # noqa: F401
        return
    except Exception:
        pass

    class _Any(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)

            def _placeholder(*a, **k):
                raise RuntimeError(f"interactive_unet.metrics.{name} is a checkpoint placeholder")
            _placeholder.__name__ = _placeholder.__qualname__ = name
            _placeholder.__module__ = "interactive_unet.metrics"
            setattr(self, name, _placeholder)
            return _placeholder

    pkg = sys.modules.setdefault("interactive_unet", types.ModuleType("interactive_unet"))
    mod = _Any("interactive_unet.metrics")
    sys.modules["interactive_unet.metrics"] = mod
    pkg.metrics = mod


def _default_loss(*args, **kwargs):
    raise RuntimeError("no loss function configured; pass loss_function= (the reference uses metrics.mcc_ce_loss)")


class UNet(_Base):
    """The UNet model (B200 engine behind the reference's interface)."""

    def __init__(self, lr=0.0001, num_channels=1, num_classes=2, loss_function=_default_loss,
                 architecture='U-Net', encoder_name='resnet34', pretrained=False):
        super().__init__()
        self.save_hyperparameters()
        self.lr = lr
        self.loss_function = loss_function
        if architecture != 'U-Net' or encoder_name not in ENCODER_BLOCKS:
            raise NotImplementedError(
                f"interactive_unet_b200 accelerates architecture='U-Net' with encoder_name in {sorted(ENCODER_BLOCKS)} "
                f"only (got {architecture!r}, {encoder_name!r}); use the stock interactive_unet.unet.UNet for others")
        if pretrained:
            raise NotImplementedError("ImageNet weights cannot be downloaded here; load a checkpoint instead")
        self.num_channels = num_channels
        self.num_classes = num_classes
        self.model = SmpUnetResnet34(num_channels, num_classes, encoder_name)     # attribute name = checkpoint key prefix
        self.softmax = nn.Softmax(dim=1)
        self._engine = None
        self._engine_key = None
        self.precision = None          # None -> engine default ("fp16", or env IU_PRECISION); or "bf16"

    # ---- engine management ----------------------------------------------------------------
    def _weights_key(self):
        return tuple((id(t), t._version) for t in self.model.state_dict(keep_vars=True).values())

    def engine(self):
        """The native engine holding the current weights (created / refreshed lazily)."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("interactive_unet_b200 runs on CUDA (sm_100a) only: move the model with "
                               ".to('cuda'); there is no CPU fallback")
        if (self._engine is None or self._engine.device != torch.device("cuda", dev.index or 0)
                or (self.precision is not None and self._engine.precision != self.precision)):
            self._engine = Engine(dev.index or 0, precision=self.precision)
            self._engine_key = None
        key = self._weights_key()
        if key != self._engine_key:
            self._engine.load_state_dict(self.model.state_dict(), self.num_classes)
            self._engine_key = key
        return self._engine

    # ---- reference interface -----------------------------------------------------------------
    def forward(self, x):
        if self.training:
            return self.softmax(self.model.forward_train(x))        # unet.py:67 (autograd path)
        if not x.is_cuda:
            raise RuntimeError("interactive_unet_b200 inference needs a CUDA input tensor; there is no CPU fallback")
        return self.engine().forward(x)

    def configure_optimizers(self):
        return torch.optim.AdamW(self.parameters(), lr=self.lr)      # unet.py:71-73

    # ---- trainer hooks (unet.py:75-116).  `metrics_module` may be injected; by default the reference's own
    #      `interactive_unet.metrics` is used (it is present wherever the reference's trainer runs).
    metrics_module = None

    def _metrics(self):
        if self.metrics_module is not None:
            return self.metrics_module
        from interactive_unet import metrics
        return metrics

    def _log_metrics(self, set_name, loss, y_hat, y, w):
        m = self._metrics()
        y, y_hat = torch.round(y), torch.round(y_hat)
        self.log(f"{set_name}/Loss", loss, prog_bar=True, on_step=False, on_epoch=True)
        for label, fn in (("Dice", m.dice), ("IoU", m.iou), ("MCC", m.mcc)):
            self.log(f"{set_name}/{label}", fn(y_hat, y, w, axes=[0, 2, 3]), on_step=False, on_epoch=True)

    def _step(self, set_name, batch):
        X, y, w = batch
        y_hat = self(X)
        loss = self.loss_function(y_hat, y, w, axes=[0, 2, 3])
        self._log_metrics(set_name, loss, y_hat, y, w)
        return loss

    def training_step(self, batch, *args):
        return self._step('train', batch)

    def validation_step(self, batch, *args):
        self._step('val', batch)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        """Lightning's classmethod re-implemented on `torch.load` (lightning is not installed here)."""
        _install_pickle_shims()
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        hp["pretrained"] = False
        allowed = ("lr", "num_channels", "num_classes", "loss_function", "architecture", "encoder_name", "pretrained")
        model = cls(**{k: v for k, v in hp.items() if k in allowed})
        model.load_state_dict(ckpt["state_dict"], strict=True)
        return model

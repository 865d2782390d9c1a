"""Seeded synthetic weights and volumes.  TEST INFRASTRUCTURE.

There is no network, so neither trained checkpoints nor datasets exist; every
test and benchmark regenerates its inputs from these seeds.
"""
import numpy as np
import torch

from .smp_unet_resnet34 import RefUNet


def noise_volume(n, seed):
    """Uniform uint8 noise `[n,n,n]` (throughput runs: the value distribution does not affect timing)."""
    return np.random.default_rng(seed).integers(0, 256, (n, n, n), dtype=np.uint8)


def blob_volume(n, seed, sigma=4.0, noise=24.0):
    """Smoothed-noise blobs plus additive noise; returns (uint8 volume, uint8 label volume)."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    field = ndimage.gaussian_filter(rng.standard_normal((n, n, n)).astype(np.float32), sigma)
    labels = (field > 0).astype(np.uint8)
    vol = 70.0 + 110.0 * labels + noise * rng.standard_normal((n, n, n)).astype(np.float32)
    return np.clip(vol, 0, 255).astype(np.uint8), labels


def _randomise_bn(model, gen):
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            c = m.num_features
            m.weight.data = 0.5 + torch.rand(c, generator=gen)
            m.bias.data = -0.2 + 0.4 * torch.rand(c, generator=gen)


@torch.no_grad()
def _calibrate_bn(model, size, gen):
    """Give BN layers the running statistics a trained net would have (activations stay O(1)),
    then jitter them so that folding is exercised with non-trivial mean/var."""
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.reset_running_stats()
        m.momentum = None
    model.train()
    for _ in range(2):
        model(torch.rand(4, 1, size, size, generator=gen))
    model.eval()
    for m in bns:
        m.momentum = 0.1
        m.running_mean += 0.1 * m.running_var.sqrt() * torch.randn(m.num_features, generator=gen)
        m.running_var *= 0.75 + 0.5 * torch.rand(m.num_features, generator=gen)


def make_model(num_classes=2, seed=1234, calib_size=64, encoder_name="resnet34"):
    """Random-init `RefUNet` in eval mode with randomised, calibrated BatchNorm statistics."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    model = RefUNet(1, num_classes, encoder_name)
    _randomise_bn(model, gen)
    _calibrate_bn(model, calib_size, gen)
    head = model.model.segmentation_head[0]
    head.bias.data = 0.1 * torch.randn(num_classes, generator=gen)
    return model.eval()


def fit_decisive(model, volume_u8, labels, steps=100, batch=8, lr=1e-3, seed=7, device="cpu"):
    """A short AdamW fit so the softmax is decisive (label agreement is meaningless on a
    random-init net whose outputs are near-ties everywhere).  Returns the model in eval mode."""
    rng = np.random.default_rng(seed)
    model = model.to(device).train()
    opt = torch.optim.AdamW(model.parameters(), lr=lr)
    n = volume_u8.shape[0]
    for _ in range(steps):
        idx = rng.integers(0, n, batch)
        axis = int(rng.integers(0, 3))
        x = np.moveaxis(volume_u8, axis, 0)[idx].astype(np.float32) / 255.0
        y = np.moveaxis(labels, axis, 0)[idx].astype(np.int64)
        x = torch.from_numpy(np.ascontiguousarray(x))[:, None].to(device)
        y = torch.from_numpy(np.ascontiguousarray(y)).to(device)
        p = model(x)
        loss = torch.nn.functional.nll_loss(torch.log(p + 1e-12), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    return model.eval()

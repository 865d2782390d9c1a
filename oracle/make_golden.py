"""Record golden vectors from the reference's own code.  TEST INFRASTRUCTURE.

Run in the build container (where `/root/reference` exists):

    python -m oracle.make_golden

It executes `/root/reference/interactive_unet/predict.py` VERBATIM
(`oracle/reference_loader.py`) and writes small fixtures to `tests/golden/`.
The reference ships no golden vectors of its own; these
files are what pins `oracle/predict_port.py` and the CUDA tail kernels on the
GPU box, where the reference tree does not exist.

The network used here is NOT a U-Net: it is `ExactToyModel`, a per-pixel,
position-dependent rational function built only from IEEE +, *, / so that the
recorded numbers are reproducible bit-for-bit on any host (no `exp`, no conv
library).  Its dependence on the pixel's (row, column) inside the slice makes
any transposition / orientation mistake in gather or accumulate visible.

`predict_volumes` (`predict.py:114-266`) is driven end to end through an
in-memory stand-in for the `zarr` module (numpy-backed arrays), so the tiling,
reflect padding, Gaussian blending and uint8 quantisation lines run exactly as
written.
"""
import glob
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

from . import reference_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class ExactToyModel(torch.nn.Module):
    """probabilities[b,c,h,w] = q_c / sum_c q_c with q_c = 0.25 + x*(c+1)/4 + ((5h + 3w + 7c) mod 11)/16."""

    def __init__(self, num_classes):
        super().__init__()
        self.num_classes = num_classes
        self.anchor = torch.nn.Parameter(torch.zeros(1))

    @property
    def device(self):
        return self.anchor.device

    def forward(self, x):
        b, _, h, w = x.shape
        hh = torch.arange(h, dtype=torch.float32)[:, None]
        ww = torch.arange(w, dtype=torch.float32)[None, :]
        q = []
        for c in range(self.num_classes):
            pos = torch.remainder(5 * hh + 3 * ww + 7 * c, 11) / 16
            q.append(0.25 + x[:, 0] * ((c + 1) / 4) + pos[None])
        q = torch.stack(q, 1)
        return q / q.sum(1, keepdim=True)


def toy_model_numpy(x, num_classes):
    """Same function as `ExactToyModel` for `[B,1,H,W]` numpy input (used by the tests)."""
    b, _, h, w = x.shape
    hh = np.arange(h, dtype=np.float32)[:, None]
    ww = np.arange(w, dtype=np.float32)[None, :]
    q = []
    for c in range(num_classes):
        pos = np.remainder(5 * hh + 3 * ww + 7 * c, 11).astype(np.float32) / np.float32(16)
        q.append(np.float32(0.25) + x[:, 0] * np.float32((c + 1) / 4) + pos[None])
    q = np.stack(q, 1).astype(np.float32)
    return q / q.sum(1, keepdims=True, dtype=np.float32)


# ---- a numpy-backed stand-in for the parts of zarr the reference calls -------------------------
class _Array:
    def __init__(self, shape, chunks, shards, dtype):
        self.data = np.zeros(shape, dtype=dtype)
        self.shape, self.chunks, self.shards, self.dtype = tuple(shape), tuple(chunks), tuple(shards), np.dtype(dtype)

    def __getitem__(self, k):
        return self.data[k].copy()

    def __setitem__(self, k, v):
        self.data[k] = v


class _Group(dict):
    def create_array(self, name, shape, chunks, shards, dtype, overwrite=True):
        self[name] = _Array(shape, chunks, shards, dtype)
        return self[name]

    def array_keys(self):
        return list(self.keys())


class _FakeZarr:
    def __init__(self):
        self.store = {}

    def open(self, path, mode="r"):
        path = str(path)
        if mode == "w":
            os.makedirs(path, exist_ok=True)
            self.store[path] = _Group()
        return self.store[path]


def run_reference_predict_volumes(volume_u8, model, input_size, num_classes, overlap=0.25, batch_size=8,
                                  axes=(0, 1, 2), chunk_size=16, shard_size=32):
    """Drive the verbatim `predict_volumes` on one in-memory volume; returns the level-0 uint8 result.

    `utils.add_multiscales` (`predict.py:261`, the pyramid post-step) is switched off here: it runs after
    level 0 is complete and cannot change it, and it raises on some 4-D shapes (it zooms the class axis
    too).  It is recorded on its own by `make_multiscales`."""
    ref = reference_loader.load()
    utils_mod = sys.modules["interactive_unet.utils"]
    unet_mod = sys.modules["interactive_unet.unet"]
    fake = _FakeZarr()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="iu_golden_")
    saved = (ref.zarr, utils_mod.zarr, unet_mod.UNet, utils_mod.add_multiscales)
    try:
        os.chdir(tmp)
        os.makedirs("data/image_volumes/vol.zarr")
        os.makedirs("data/predicted_volumes")
        g = fake.open("data/image_volumes/vol.zarr", mode="w")
        arr = g.create_array("0", volume_u8.shape, (chunk_size,) * 3, (shard_size,) * 3, "uint8")
        arr[:] = volume_u8
        ref.zarr = fake
        utils_mod.zarr = fake
        unet_mod.UNet = lambda **kw: model          # predict.py:134 (no checkpoint on disk)
        utils_mod.add_multiscales = lambda *a, **k: None
        ref.predict_volumes(input_size=input_size, num_channels=1, num_classes=num_classes, overlap=overlap,
                            chunk_size=chunk_size, shard_size=shard_size, batch_size=batch_size, axes=list(axes))
        return fake.store["data/predicted_volumes/vol.zarr"]["0"].data.copy()
    finally:
        ref.zarr, utils_mod.zarr, unet_mod.UNet, utils_mod.add_multiscales = saved
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference_add_multiscales(volume, chunks, shards, scale=0.5):
    """Drive the verbatim `utils.add_multiscales` (`utils.py:50-80`) on one in-memory store; returns the levels it
    created as {name: ndarray}, and the exception it ended with (None, or (type name, message))."""
    reference_loader.load()
    utils_mod = sys.modules["interactive_unet.utils"]
    fake = _FakeZarr()
    saved = utils_mod.zarr
    tmp = tempfile.mkdtemp(prefix="iu_golden_")
    try:
        utils_mod.zarr = fake
        g = fake.open(os.path.join(tmp, "x.zarr"), mode="w")
        arr = g.create_array("0", volume.shape, chunks, shards, volume.dtype)
        arr[:] = volume
        err = None
        try:
            utils_mod.add_multiscales(os.path.join(tmp, "x.zarr"), scale=scale)
        except Exception as e:                      # the reference's own failure modes are part of the record
            err = (type(e).__name__, str(e))
        return {k: v.data.copy() for k, v in g.items() if k != "0"}, err
    finally:
        utils_mod.zarr = saved
        shutil.rmtree(tmp, ignore_errors=True)


MULTISCALE_CASES = [
    # name, shape, inner chunk edge, shard edge
    ("c2", (32, 32, 32, 2), 8, 16),             # class axis 2 -> 1 -> 0
    ("c4", (32, 32, 32, 4), 8, 16),             # class axis 4 -> 2 -> 1 (keeps classes 0 and 3)
    ("ragged_c2", (48, 20, 36, 2), 4, 8),       # several levels, edge blocks shorter than a shard
    ("image", (40, 36, 44), 16, 32),            # 3-D image volume (create_multiscale_zarr)
    ("fill_c2", (64, 32, 32, 2), 16, 32),       # 32 -> 16 blocks: scipy's constant fill on the last plane of a block
    ("u16", (24, 32, 16), 4, 8),                # 16-bit image volume
    ("odd", (38, 32, 32), 8, 16),               # ValueError at level 2 (19 -> 9)
    ("c3", (32, 16, 16, 3), 8, 16),             # ValueError at level 1 (class axis 3: round(1.5) = 2 vs int(1.5) = 1)
    ("small", (8, 8, 8, 2), 8, 16),             # fits a chunk: UnboundLocalError in the reference (utils.py:77)
]


def make_multiscales():
    """6. utils.add_multiscales / resize_volume (utils.py:29-80): the pyramid, incl. the zoomed class axis."""
    rng = np.random.default_rng(20261019)
    rec = {}
    for name, shape, chunk, shard in MULTISCALE_CASES:
        dtype = np.uint16 if name == "u16" else np.uint8
        vol = rng.integers(1, np.iinfo(dtype).max, shape, dtype=dtype)     # no zeros: scipy's fill voxels stand out
        tail = shape[3:]
        levels, err = run_reference_add_multiscales(vol, (chunk,) * 3 + tail, (shard,) * 3 + tail)
        rec[f"{name}_volume"] = vol
        rec[f"{name}_grid"] = np.array([chunk, shard])
        rec[f"{name}_levels"] = np.array(sorted(int(k) for k in levels))
        for k, v in levels.items():
            rec[f"{name}_level{k}"] = v
        rec[f"{name}_error"] = np.array(list(err) if err else ["", ""])
        print(name, shape, {k: v.shape for k, v in levels.items()}, err)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "multiscales.npz"), **rec)


def main():
    ref = reference_loader.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # 1. predict_block (predict.py:79-112): orientation + accumulate order + /len(axes)
    for name, size, classes, axes, bs in [("block_s16_c2_a012", 16, 2, [0, 1, 2], 8),
                                          ("block_s16_c4_a012", 16, 4, [0, 1, 2], 16),
                                          ("block_s16_c3_a20", 16, 3, [2, 0], 4),
                                          ("block_s8_c2_a1", 8, 2, [1], 8)]:
        vol = rng.integers(0, 256, (size,) * 3, dtype=np.uint8)
        block = torch.tensor(vol.astype("float32") / 255.0)          # predict.py:237
        out = ref.predict_block(ExactToyModel(classes), block, num_classes=classes, batch_size=bs, axes=axes)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), volume=vol, axes=np.array(axes),
                            batch_size=bs, num_classes=classes, mean_probs=out)

    # 2. gaussian_3d (predict.py:327-347)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gaussian3d.npz"),
                        **{f"w{s}": ref.gaussian_3d(s, sigma=0.125) for s in (8, 16, 32, 48)},
                        **{f"diag{s}": np.stack([ref.gaussian_3d(s)[i, i, i] for i in range(s)]) for s in (128,)},
                        **{f"row{s}": ref.gaussian_3d(s)[s // 2 - 1, 3, :] for s in (128,)})

    # 3. get_block_coordinates (predict.py:362-411), get_shard_coordinates (:318-325)
    cases = [((64, 64, 64), 64, 0.25), ((128, 128, 128), 64, 0.25), ((100, 80, 60), 32, 0.25),
             ((70, 70, 70), 64, 0.25), ((512, 512, 512), 256, 0.25), ((40, 36, 44), 32, 0.25),
             ((96, 96, 96), 32, 0.5)]
    rec = {}
    for n, (shape, size, ov) in enumerate(cases):
        c, p, l = ref.get_block_coordinates(np.array(shape), input_size=size, overlap=ov)
        rec[f"case{n}_args"] = np.array(list(shape) + [size, int(ov * 100)])
        rec[f"case{n}_clipped"], rec[f"case{n}_padded"], rec[f"case{n}_local"] = c, p, l
    rec["shards_100_80_60_32"] = ref.get_shard_coordinates(np.array((100, 80, 60)), shard_size=32)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "coordinates.npz"), **rec)

    # 4. get_padded_block (predict.py:291-316)
    vol = rng.integers(0, 256, (12, 10, 14), dtype=np.uint8)
    boxes = np.array([[-3, -2, -4, 9, 8, 10], [2, 1, 3, 10, 9, 11], [4, 3, 5, 15, 12, 17], [-2, 0, 6, 6, 8, 16]])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "padded_block.npz"), volume=vol, boxes=boxes,
                        **{f"out{i}": ref.get_padded_block(vol, *b) for i, b in enumerate(boxes)})

    # 5. predict_volumes end to end (predict.py:114-266): single block and tiled
    for name, shape, size, classes, axes in [("volume_single_s32_c2", (32, 32, 32), 32, 2, [0, 1, 2]),
                                             ("volume_single_s32_c4", (32, 32, 32), 32, 4, [0, 1, 2]),
                                             ("volume_tiled_s32_c2", (40, 36, 44), 32, 2, [0, 1, 2]),
                                             ("volume_tiled_s16_c3", (40, 24, 33), 16, 3, [0, 2])]:
        vol = rng.integers(0, 256, shape, dtype=np.uint8)
        out = run_reference_predict_volumes(vol, ExactToyModel(classes), size, classes, axes=axes)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), volume=vol, input_size=size,
                            num_classes=classes, axes=np.array(axes), out_u8=out)

    make_multiscales()

    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        print(f"{os.path.getsize(f):8d}  {os.path.relpath(f)}")


if __name__ == "__main__":
    main()

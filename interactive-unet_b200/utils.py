"""Drop-in for the volume-store helpers of `/root/reference/interactive_unet/utils.py:18-98`:
`read_volume`, `resize_volume`, `add_multiscales`, `create_multiscale_zarr`.

The stores are Zarr v3 with sharding, read and written by `zarr3.py`; the voxel work -- the nearest-neighbour 0.5x
zoom of every pyramid level and the re-ordering between `[D,H,W,C]` and the store's inner chunks -- runs on the GPU
through `libiunet_b200.so` (`iu_engine_zoom_nearest`, `iu_engine_to_chunks`, `iu_engine_from_chunks`); host threads
only (de)compress whole chunks.  There is no CPU path for the voxel work.

What the reference computes, restated exactly (`zoom_tables`):
`resize_volume` cuts the source into blocks of `block_size` along the first three axes and assigns
`scipy.ndimage.zoom(block, scale, order=0)` into `dst[int(i0*scale):int(i1*scale), ...]`.  For `order=0` (the only
value the reference uses) `zoom` is separable: output sample `o` of an axis with `n` samples zoomed to
`m = round(n*scale)` takes input sample `floor(o*(n-1)/(m-1) + 0.5)` (float64), and is the constant fill 0 where
`o*(n-1)/(m-1) > n-1` through rounding (scipy's `mode='constant'`; it happens e.g. for n = 32 -> 16).  A trailing class
axis is zoomed too (4-D blocks): C = 2 -> 1 keeps class 0, C = 4 -> 2 keeps classes 0 and 3, C = 1 -> an empty level.
Shapes where numpy cannot assign the zoomed block (`round(n*scale) != int(i1*scale) - int(i0*scale)`, e.g. odd
extents, C = 3) raise the same `ValueError` as the reference.
"""
import threading

import numpy as np
import torch

from . import zarr3
from .engine import Engine

__all__ = ["read_volume", "resize_volume", "add_multiscales", "create_multiscale_zarr", "zoom_tables",
           "write_array_from_device", "read_array_to_device", "stage_array"]

_engines = {}


def _engine(device=None):
    """A weight-less engine for the staging kernels (one per device)."""
    if not torch.cuda.is_available():
        raise RuntimeError("interactive_unet_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(idx)
    return _engines[idx]


def read_volume(path, level=0):
    """`utils.py:18-27`."""
    root = zarr3.open(path, mode='r')
    num_scales = len(np.sort(list(root.array_keys())))
    level = int(np.clip(level, 0, num_scales))
    return root[str(level)]


# ------------------------------------------------------------------------------------------------- zoom tables
def _zoom_axis(n, scale):
    """scipy.ndimage.zoom(order=0, mode='constant', grid_mode=False) along one axis of `n` samples: source index per
    output sample, -1 where scipy writes the constant 0."""
    m = int(round(n * scale))
    if m <= 0:
        return np.zeros(0, np.int64)
    z = (n - 1) / (m - 1) if m > 1 else 1.0
    cc = np.arange(m, dtype=np.float64) * z
    idx = np.floor(cc + 0.5).astype(np.int64)
    idx[(cc < 0) | (cc > n - 1)] = -1
    return idx


def zoom_tables(src_shape, dst_shape, scale=0.5, block_size=512):
    """Per-axis gather tables equivalent to `resize_volume(src, dst, scale, block_size, order=0)` (`utils.py:29-48`):
    `dst[i,j,k,...] = src[t0[i], t1[j], t2[k], ...]` (0 where a table holds -1).  Raises the reference's `ValueError`
    (numpy's broadcast message, for the first block in the reference's loop order that cannot be assigned)."""
    src_shape = tuple(int(v) for v in src_shape)
    dst_shape = tuple(int(v) for v in dst_shape)
    if len(src_shape) < 3 or len(dst_shape) != len(src_shape):
        raise ValueError("resize_volume works on arrays with at least three axes")
    block_size = int(block_size)
    tables, per_axis = [], []
    for ax in range(3):
        n, table, blocks = src_shape[ax], np.full(dst_shape[ax], -1, np.int64), []
        for i0 in range(0, n, block_size):
            i1 = min(i0 + block_size, n)
            t0, t1 = int(i0 * scale), int(i1 * scale)
            lo, hi = min(t0, dst_shape[ax]), min(t1, dst_shape[ax])            # numpy clips the target slice
            idx = _zoom_axis(i1 - i0, scale)
            blocks.append((idx.size, max(hi - lo, 0)))
            if idx.size == hi - lo or idx.size == 1:                           # equal, or broadcast of one sample
                table[lo:hi] = np.where(idx >= 0, idx + i0, -1)
        tables.append(table)
        per_axis.append(blocks)
    trailing_val, trailing_dst = [], []
    for ax in range(3, len(src_shape)):                                        # trailing axes are zoomed whole
        idx = _zoom_axis(src_shape[ax], scale)
        trailing_val.append(idx.size)
        trailing_dst.append(dst_shape[ax])
        t = np.full(dst_shape[ax], -1, np.int64)
        if idx.size == dst_shape[ax] or idx.size == 1:
            t[:] = idx
        tables.append(t)

    def assignable(v, d):
        return v == d or v == 1
    for bi in per_axis[0]:
        for bj in per_axis[1]:
            for bk in per_axis[2]:
                val = [bi[0], bj[0], bk[0]] + trailing_val
                dst = [bi[1], bj[1], bk[1]] + trailing_dst
                if not all(assignable(v, d) for v, d in zip(val, dst)):
                    fmt = lambda s: "(" + ",".join(str(v) for v in s) + ("," if len(s) == 1 else "") + ")"
                    raise ValueError(f"could not broadcast input array from shape {fmt(val)} into shape {fmt(dst)}")
    return [t.astype(np.int32) for t in tables]


# ------------------------------------------------------------------------------------------------- device <-> store
def _chunked_on_three_axes(arr):
    """The bulk path's precondition: one chunk spans the trailing (class) extent -- exactly, or with padding when a
    pyramid level inherits level 0's chunk shape (`utils.py:66-71`) while its class axis was halved."""
    if arr.ndim < 3 or arr.chunk_grid[3:] != (1,) * (arr.ndim - 3):
        return False
    return tuple(arr.chunks[3:]) == tuple(arr.shape[3:]) or arr.ndim == 4


class _PinnedPool:
    """Page-locked staging buffers, kept between calls: `cudaHostAlloc` of a few hundred MiB costs more than the copy
    it serves (measured 0.1 s for 256 MiB), and `predict_volumes` needs the same sizes for every volume."""

    def __init__(self):
        self._free = []
        self._lock = threading.Lock()

    def acquire(self, shape, dtype):
        nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size()
        with self._lock:
            fits = [k for k, b in enumerate(self._free) if b.numel() >= nbytes]
            buf = self._free.pop(min(fits, key=lambda k: self._free[k].numel())) if fits else None
        if buf is None:
            buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        return buf, buf[:nbytes].view(dtype).view(*shape)

    def release(self, buf):
        with self._lock:
            self._free.append(buf)
            self._free.sort(key=lambda b: b.numel())
            while sum(b.numel() for b in self._free) > (8 << 30):        # keep at most 8 GiB parked
                self._free.pop(0)


_pinned = _PinnedPool()


class PendingWrite:
    """A store write whose compression / file tasks are still running in the pool (`wait=False`)."""

    def __init__(self, futures, buf):
        self._futures, self._buf = futures, buf

    def result(self):
        try:
            for f in self._futures:
                f.result()
        finally:
            if self._buf is not None:
                _pinned.release(self._buf)
                self._buf = None


class StagedArray:
    """A store array decoded into pinned host memory in chunk-major order (`stage_array`), waiting for its upload."""

    def __init__(self, arr, buf, staged):
        self.arr, self._buf, self._staged = arr, buf, staged

    def to_device(self, device=None):
        """One H2D copy + `iu_engine_from_chunks`; the pinned buffer goes back to the pool."""
        eng = _engine(device)
        try:
            staged_dev = self._staged.to(eng.device, non_blocking=True)
            torch.cuda.current_stream(eng.device).synchronize()
        finally:
            self.release()
        return eng.from_chunks(staged_dev, self.arr.shape, self.arr.chunks)

    def release(self):
        if self._buf is not None:
            _pinned.release(self._buf)
            self._buf = self._staged = None


def stage_array(arr):
    """Host half of `read_array_to_device` for a store array on the bulk path: pool threads decompress the inner
    chunks into a pinned buffer.  Touches no device state, so it can run ahead on another thread (the next volume is
    staged while the current one is being predicted, `predict_volumes`)."""
    if not isinstance(arr, zarr3.Array) or arr.size == 0 or not _chunked_on_three_axes(arr):
        raise ValueError("stage_array: a non-empty zarr3.Array chunked over its first three axes is required")
    buf, staged = _pinned.acquire(arr.chunk_major_shape(), torch.from_numpy(np.empty(0, arr.dtype)).dtype)
    try:
        arr.read_chunk_major(out=staged.numpy())
    except BaseException:
        _pinned.release(buf)
        raise
    return StagedArray(arr, buf, staged)


def read_array_to_device(arr, device=None):
    """Whole `zarr3.Array` (or any array-like) -> CUDA tensor of its shape.  Store arrays chunked over the first three
    axes take the bulk path: host threads decompress inner chunks into pinned memory, one H2D copy, and
    `iu_engine_from_chunks` puts the voxels in `[D,H,W,...]` order."""
    eng = _engine(device)
    if not isinstance(arr, zarr3.Array):
        return torch.as_tensor(np.ascontiguousarray(arr)).to(eng.device)
    if arr.size == 0:
        return torch.empty(arr.shape, dtype=torch.from_numpy(np.empty(0, arr.dtype)).dtype, device=eng.device)
    if not _chunked_on_three_axes(arr):
        return torch.from_numpy(arr[...]).to(eng.device)
    return stage_array(arr).to_device(eng.device)


def write_array_from_device(arr, data, wait=True):
    """CUDA tensor of `arr.shape` -> the whole `zarr3.Array`: `iu_engine_to_chunks` on the device, one D2H copy into
    pinned memory, host threads compress the inner chunks and write one shard file each.  `wait=False` returns a
    `PendingWrite` as soon as the data has left the device; call its `result()` before relying on the files."""
    if tuple(data.shape) != tuple(arr.shape):
        raise ValueError(f"could not broadcast input array from shape {tuple(data.shape)} into shape {tuple(arr.shape)}")
    if arr.size == 0:
        return PendingWrite([], None)
    if not _chunked_on_three_axes(arr):
        arr[...] = data.cpu().numpy()
        return PendingWrite([], None)
    eng = _engine(data.device)
    staged_dev = eng.to_chunks(data.contiguous(), arr.chunks)
    buf, staged = _pinned.acquire(staged_dev.shape, staged_dev.dtype)
    staged.copy_(staged_dev)
    del staged_dev
    pending = PendingWrite(arr.write_chunk_major(staged.numpy(), wait=False), buf)
    if wait:
        pending.result()
    return pending


def resize_volume(src_vol, dst_vol, scale=0.5, block_size=512, order=0, _pending=None):
    """`utils.py:29-48` for `order=0` (the reference's only use, `utils.py:74`): `dst_vol` <- block-wise nearest zoom of
    `src_vol`.  `src_vol` may be a `zarr3.Array`, a numpy array or a CUDA tensor; `dst_vol` a `zarr3.Array` or a CUDA
    tensor of the target shape.  Returns the zoomed level as a CUDA tensor (so a pyramid never re-reads the store)."""
    if order != 0:
        raise NotImplementedError("resize_volume: only order=0 (nearest), the reference's own setting, is implemented")
    tables = zoom_tables(src_vol.shape, dst_vol.shape, scale=scale, block_size=block_size)
    src = src_vol if isinstance(src_vol, torch.Tensor) and src_vol.is_cuda else read_array_to_device(src_vol)
    eng = _engine(src.device)
    out = eng.zoom_nearest(src.contiguous(), tables, out=dst_vol if isinstance(dst_vol, torch.Tensor) else None)
    if isinstance(dst_vol, zarr3.Array):
        pending = write_array_from_device(dst_vol, out, wait=_pending is None)
        if _pending is not None:
            _pending.append(pending)
    elif not isinstance(dst_vol, torch.Tensor):
        dst_vol[...] = out.cpu().numpy()
    return out


def _num_steps(volume_shape, chunk_shape, scale):
    """`utils.py:60`: downscale steps until the volume fits inside a chunk."""
    return int(np.floor(np.log((np.array(volume_shape) / np.array(chunk_shape)).max()) / np.log(1 / scale)))


def add_multiscales(src_file, scale=0.5, level0=None, _defer=None):
    """`utils.py:50-80`: levels '1', '2', ... of `src_file`, each the block-wise nearest 0.5x zoom of the previous one
    (block = one shard), until the volume fits a chunk.  `level0`: the level-'0' data as a CUDA tensor if the caller
    still holds it (saves reading the store back).

    One deliberate difference: when no level is needed (the volume already fits a chunk) the reference ends in an
    `UnboundLocalError` at its `del root, z0, z1` (`utils.py:77`) after doing nothing; this returns normally."""
    root = zarr3.open(src_file, mode='r+')
    z0 = root['0']
    volume_shape, chunk_shape, shard_shape = z0.shape, z0.chunks, z0.shards
    if shard_shape is None:
        raise ValueError(f"{src_file}: level '0' is not sharded (the reference creates every level with shards=)")
    num_steps = _num_steps(volume_shape, chunk_shape, scale)
    cur = level0
    pending = []              # every level's compression runs in the pool while the next level is zoomed on the device
    try:
        for i in range(num_steps):
            z0 = root[str(i)]
            z1_shape = tuple(int(x * scale) for x in z0.shape)
            z1 = root.create_array(name=str(i + 1), shape=z1_shape, chunks=chunk_shape, shards=shard_shape,
                                   dtype=z0.dtype, overwrite=True)
            cur = resize_volume(cur if cur is not None else z0, z1, scale=scale, block_size=shard_shape[0], order=0,
                                _pending=pending)
    finally:
        if _defer is not None:
            _defer.extend(pending)     # the caller waits (predict_volumes overlaps them with the next volume)
        else:
            for p in pending:
                p.result()


def create_multiscale_zarr(volume, dst_file, scale=0.5, chunk_size=128, shard_size=256):
    """`utils.py:82-98`: a new store with `volume` as level '0' and its pyramid."""
    chunk_shape = (chunk_size, chunk_size, chunk_size)
    shard_shape = (shard_size, shard_size, shard_size)
    root = zarr3.open(dst_file, mode='w')
    z0 = root.create_array(name='0', shape=volume.shape, chunks=chunk_shape, shards=shard_shape,
                           dtype=volume.dtype if not isinstance(volume, torch.Tensor) else
                           torch.empty(0, dtype=volume.dtype).numpy().dtype, overwrite=True)
    dev = volume if isinstance(volume, torch.Tensor) and volume.is_cuda else \
        torch.as_tensor(np.ascontiguousarray(volume)).to(_engine().device)
    write_array_from_device(z0, dev)
    add_multiscales(dst_file, scale=scale, level0=dev)

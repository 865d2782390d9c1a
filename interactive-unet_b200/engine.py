"""Python handle on the native engine (`libiunet_b200.so`).

Thin by design: argument marshalling, pointer extraction from numpy arrays (host) and torch CUDA
tensors (device), and error mapping.  All arithmetic happens in the library's sm_100a kernels.
"""
import ctypes
import os
import threading

import numpy as np
import torch

from . import _lib

STATE_DICT_PREFIX = "model."          # attribute name at `unet.py:56` of the reference


def gaussian_window_1d(size, sigma=0.125, eps=1e-3):
    """The 1-D factor, global maximum and lower clip bound of the reference's `gaussian_3d`
    (`predict.py:327-347`), from which the reduce kernel rebuilds the 3-D window per voxel.

    Returns (g fp32[size], gmax, lo) with window[z,y,x] = clip((g[z]*g[y])*g[x] / gmax, lo, 1)."""
    s = sigma * size
    coords = np.arange(size, dtype=np.float32) - (size - 1) / 2.0
    g = np.exp(-(coords ** 2) / (2 * s ** 2)).astype(np.float32)
    g /= g.max()
    gm = g.max()
    gmax = np.float32(np.float32(gm * gm) * gm)
    gmin = np.float32(np.float32(np.float32(g.min() * g.min()) * g.min()) / gmax)
    lo = np.float32(max(gmin, eps))
    return np.ascontiguousarray(g), float(gmax), float(lo)


def _ptr(a):
    """(pointer, keep-alive object) of a numpy array or torch tensor; None -> NULL."""
    if a is None:
        return None, None
    if isinstance(a, torch.Tensor):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return ctypes.c_void_p(a.data_ptr()), a
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return ctypes.c_void_p(a.ctypes.data), a
    raise TypeError(f"unsupported buffer type {type(a)}")


def _dtype_code(a):
    dt = a.dtype
    if dt in (torch.uint8, np.dtype("uint8")):
        return _lib.DTYPE_U8
    if dt in (torch.float32, np.dtype("float32")):
        return _lib.DTYPE_F32
    raise TypeError(f"volume dtype must be uint8 or float32, got {dt}")


class _BatchLimit:
    def __init__(self, engine, max_batch):
        self.engine, self.max_batch = engine, max_batch

    def __enter__(self):
        self.engine._lock.acquire()
        try:
            self.engine.set_max_batch(self.max_batch or 0)
        except BaseException:
            self.engine._lock.release()
            raise
        return self.engine

    def __exit__(self, *exc):
        try:
            self.engine.set_max_batch(0)
        finally:
            self.engine._lock.release()


class Engine:
    """One native engine bound to one CUDA device.  Calls are serialised with a lock because
    the reference calls the prediction entry points from worker threads (`app.py:737-739`)."""

    def __init__(self, device=0, precision=None):
        """`precision`: 16-bit storage format of weights / activations, "fp16" (default) or "bf16"
        (env `IU_PRECISION` overrides the default).  Accumulation is fp32 either way."""
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self._lock = threading.RLock()
        if isinstance(device, torch.device):
            device = device.index or 0
        precision = precision or os.environ.get("IU_PRECISION", "fp16")
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        rc = self._lib.iu_engine_create(int(device), ctypes.byref(self._h))
        if rc != _lib.IU_OK:
            msg = self._lib.iu_last_error(None)
            raise _lib.EngineError(rc, msg.decode() if msg else "iu_engine_create failed")
        self._check(self._lib.iu_engine_set_precision(self._h, _lib.PRECISIONS[precision]))
        self.precision = precision
        self.act_dtype = torch.float16 if precision == "fp16" else torch.bfloat16
        self.device = torch.device("cuda", int(device))
        self.num_classes = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.iu_engine_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        _lib.check(self._lib, self._h, rc)

    def _sync_torch(self, *bufs):
        # device buffers produced on torch's current stream must be complete before our stream reads them
        if any(isinstance(b, torch.Tensor) and b.is_cuda for b in bufs):
            torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, num_classes):
        """Upload smp.Unet('resnet34') weights given with the reference's key names (`model.` prefix optional)."""
        names, arrays = [], []
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked"):
                continue
            if k.startswith(STATE_DICT_PREFIX):
                k = k[len(STATE_DICT_PREFIX):]
            a = v.detach().to("cpu", torch.float32).contiguous().numpy() if isinstance(v, torch.Tensor) \
                else np.ascontiguousarray(v, dtype=np.float32)
            names.append(k.encode())
            arrays.append(a)
        n = len(names)
        c_names = (ctypes.c_char_p * n)(*names)
        c_data = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrays])
        c_numel = (ctypes.c_int64 * n)(*[a.size for a in arrays])
        with self._lock:
            self._check(self._lib.iu_engine_load_weights(self._h, int(num_classes), n, c_names, c_data, c_numel))
        self.num_classes = int(num_classes)

    def set_max_batch(self, max_batch):
        self._check(self._lib.iu_engine_set_max_batch(self._h, int(max_batch or 0)))

    def limit_batch(self, max_batch):
        """Context manager: hold the engine for the caller's thread with `max_batch` as the cap on slices per network
        pass, and lift the cap on exit.  The app calls `predict_slice` and `predict_volumes` from different worker
        threads on one cached model (`app.py:737-739`); set / predict / reset must not interleave between them."""
        return _BatchLimit(self, max_batch)

    def auto_batch(self, h, w, count):
        """Slices per network pass the engine would choose for `count` slices of h x w."""
        return int(self._lib.iu_engine_auto_batch(self._h, int(h), int(w), int(count)))

    def release_workspace(self):
        """Give activation plans and pooled scratch back to the driver (weights stay)."""
        with self._lock:
            self._check(self._lib.iu_engine_release_workspace(self._h))

    def held_bytes(self):
        return int(self._lib.iu_engine_held_bytes(self._h))

    # ---- stream ordering against torch (no host synchronisation)
    def _ext_stream(self):
        if getattr(self, "_ext", None) is None:
            self._ext = torch.cuda.ExternalStream(self.stream_handle(), device=self.device)
        return self._ext

    def wait_torch(self):
        """The engine's stream waits for everything queued so far on torch's current stream."""
        self._ext_stream().wait_event(torch.cuda.current_stream(self.device).record_event())

    def torch_wait(self):
        """torch's current stream waits for everything queued so far on the engine's stream."""
        torch.cuda.current_stream(self.device).wait_event(self._ext_stream().record_event())

    def workspace_bytes(self, batch, h, w):
        return int(self._lib.iu_engine_workspace_bytes(self._h, batch, h, w))

    def launch_count(self):
        return int(self._lib.iu_engine_launch_count(self._h))

    def profile(self, enable=True):
        """Bracket every kernel launch with CUDA events on the engine's stream (per-class timing)."""
        self._check(self._lib.iu_engine_profile(self._h, int(bool(enable))))

    def profile_read(self, reset=True):
        """dict class -> (kernel milliseconds, launches) accumulated while profiling was enabled."""
        n = len(_lib.PROF_CLASSES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        self._check(self._lib.iu_engine_profile_read(self._h, ms, cnt, int(bool(reset))))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(_lib.PROF_CLASSES)}

    def debug_counters(self, n_layers=48, reset=True):
        """Per-layer role cycle counters of the halo kernel (needs env IU_CONV_DEBUG=1 at engine creation)."""
        buf = (ctypes.c_uint64 * (16 * n_layers))()
        self._check(self._lib.iu_engine_debug_counters(self._h, buf, 16 * n_layers, int(bool(reset))))
        return np.array(buf, dtype=np.uint64).reshape(n_layers, 16)

    def stream_handle(self):
        return int(self._lib.iu_engine_stream(self._h) or 0)

    def synchronize(self):
        self._check(self._lib.iu_engine_synchronize(self._h))

    # ------------------------------------------------------------------ network
    def forward(self, x, out=None):
        """`UNet.forward`: fp32 [B,1,H,W] -> probabilities fp32 [B,C,H,W]; numpy in -> numpy out,
        CUDA tensor in -> CUDA tensor out."""
        b, ch, h, w = x.shape
        if ch != 1:
            raise ValueError("the engine supports num_channels=1 only")
        if isinstance(x, torch.Tensor):
            x = x.to(torch.float32).contiguous()
            if out is None:
                out = torch.empty((b, self.num_classes, h, w), dtype=torch.float32, device=x.device)
        else:
            x = np.ascontiguousarray(x, dtype=np.float32)
            if out is None:
                out = np.empty((b, self.num_classes, h, w), dtype=np.float32)
        xp, _ = _ptr(x)
        op, _ = _ptr(out)
        with self._lock:
            self._sync_torch(x, out)
            self._check(self._lib.iu_engine_forward(self._h, xp, b, h, w, op, 0))
        return out

    # ------------------------------------------------------------------ volume path
    def gather_slices(self, volume, axis, start, count):
        n = volume.shape[0]
        out = torch.empty((count, n, n), dtype=torch.float32, device=self.device)
        vp, _ = _ptr(volume)
        with self._lock:
            self._sync_torch(volume, out)
            self._check(self._lib.iu_engine_gather_slices(self._h, vp, _dtype_code(volume), n, int(axis), int(start),
                                                          int(count), ctypes.c_void_p(out.data_ptr()), 0))
        return out

    def predict_axis(self, volume, axis, slice_begin=0, slice_count=None, out=None, slice_offset=0,
                     slice_total=None, row_block=None, asynchronous=False):
        """Probabilities of the slices [slice_begin, slice_begin+slice_count) along `axis` (device fp32)."""
        n = volume.shape[0]
        slice_count = n - slice_begin if slice_count is None else slice_count
        slice_total = slice_count if slice_total is None else slice_total
        row_block = n if row_block is None else row_block
        if out is None:
            out = torch.empty((slice_total, n, n, self.num_classes), dtype=torch.float32, device=self.device)
        vp, _ = _ptr(volume)
        with self._lock:
            self._sync_torch(volume, out)
            self._check(self._lib.iu_engine_predict_axis(
                self._h, vp, _dtype_code(volume), n, int(axis), int(slice_begin), int(slice_count),
                ctypes.c_void_p(out.data_ptr()), int(slice_offset), int(slice_total), int(row_block),
                _lib.FLAG_ASYNC if asynchronous else 0))
        return out

    def predict_slices(self, source, offset, count, h, w, strides, out, slice_offset=0, slice_total=None,
                       row_block=None, asynchronous=False, sync=True):
        """Probabilities of `count` slices of h x w read from a CUDA tensor through element strides:
        pixel (slice i, row r, col c) = source.view(-1)[offset + i*strides[0] + r*strides[1] + c*strides[2]].
        `sync=False`: the caller orders the streams itself (`wait_torch` / `torch_wait`)."""
        slice_total = count if slice_total is None else slice_total
        row_block = h if row_block is None else row_block
        base = ctypes.c_void_p(source.data_ptr() + int(offset) * source.element_size())
        with self._lock:
            if sync:
                self._sync_torch(source, out)
            self._check(self._lib.iu_engine_predict_slices(
                self._h, base, _dtype_code(source), int(count), int(h), int(w), int(strides[0]), int(strides[1]),
                int(strides[2]), ctypes.c_void_p(out.data_ptr()), int(slice_offset), int(slice_total), int(row_block),
                _lib.FLAG_ASYNC if asynchronous else 0))
        return out

    def reduce(self, probs, order, n, t=None, z0=0, window=None, out_u8=None, out_labels=None, out_mean=None,
               asynchronous=False, zoff=0, zcount=None, sync=True):
        """K4.  `probs`: dict axis -> device fp32 tensor; `order`: accumulation order (`predict.py:87`);
        `window`: None or (g1d, gmax, lo) from `gaussian_window_1d`; `zoff` / `zcount`: reduce only these planes of the
        slab; `sync=False`: the caller orders the streams itself."""
        t = n if t is None else t
        zcount = t - zoff if zcount is None else zcount
        ptrs = [ctypes.c_void_p(probs[a].data_ptr()) if a in probs and probs[a] is not None else None
                for a in (0, 1, 2)]
        c_order = (ctypes.c_int * len(order))(*[int(a) for a in order])
        g, gmax, lo = (None, 1.0, 0.0) if window is None else window
        gp, _g = _ptr(g)
        with self._lock:
            if sync:
                self._sync_torch(*[p for p in probs.values() if p is not None], out_u8, out_labels, out_mean)
            self._check(self._lib.iu_engine_reduce_planes(
                self._h, ptrs[0], ptrs[1], ptrs[2], c_order, len(order), int(n), int(t), int(z0), int(zoff),
                int(zcount), int(self.num_classes), gp, float(gmax), float(lo), _ptr(out_u8)[0], _ptr(out_labels)[0],
                _ptr(out_mean)[0], _lib.FLAG_ASYNC if asynchronous else 0))

    def predict_volume(self, volume, axes=(0, 1, 2), window=None, out_u8=None, out_labels=None, out_mean=None):
        """Whole single-GPU path; every buffer may live on the host (numpy) or on the device (torch)."""
        n = volume.shape[0]
        if tuple(volume.shape) != (n, n, n):
            raise ValueError("the engine predicts cubic blocks only (predict.py:81)")
        c_axes = (ctypes.c_int * len(axes))(*[int(a) for a in axes])
        g, gmax, lo = (None, 1.0, 0.0) if window is None else window
        vp, _ = _ptr(volume)
        gp, _g = _ptr(g)
        with self._lock:
            self._sync_torch(volume, out_u8, out_labels, out_mean)
            self._check(self._lib.iu_engine_predict_volume(
                self._h, vp, _dtype_code(volume), n, c_axes, len(axes), gp, float(gmax), float(lo),
                _ptr(out_u8)[0], _ptr(out_labels)[0], _ptr(out_mean)[0], 0))

    # ------------------------------------------------------------------ tiled / blended mode (predict.py:201,235-256)
    def predict_tiled(self, volume, input_size, origins, axes=(0, 1, 2), window=None, out_u8=None, out_labels=None):
        """uint8 volume `[D,H,W]` (numpy or CUDA tensor) predicted in cubic blocks of edge `input_size` whose voxel
        (0,0,0) sits at `origins[b]` (the first three columns of `padded_block_coords`); Gaussian-blended, quantised."""
        d, h, w = (int(v) for v in volume.shape)
        org = np.ascontiguousarray(np.asarray(origins, dtype=np.int32).reshape(-1, 3))
        c_axes = (ctypes.c_int * len(axes))(*[int(a) for a in axes])
        g, gmax, lo = window if window is not None else gaussian_window_1d(input_size)
        vp, _ = _ptr(volume)
        gp, _g = _ptr(g)
        with self._lock:
            self._sync_torch(volume, out_u8, out_labels)
            self._check(self._lib.iu_engine_predict_tiled(
                self._h, vp, d, h, w, int(input_size), int(org.shape[0]), org.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                c_axes, len(axes), gp, float(gmax), float(lo), _ptr(out_u8)[0], _ptr(out_labels)[0], 0))

    def extract_block(self, volume, origin, size):
        """`get_padded_block` on a CUDA uint8 volume -> CUDA uint8 `[size,size,size]`."""
        d, h, w = (int(v) for v in volume.shape)
        out = torch.empty((size, size, size), dtype=torch.uint8, device=volume.device)
        with self._lock:
            self._sync_torch(volume)
            self._check(self._lib.iu_engine_extract_block(self._h, _ptr(volume)[0], d, h, w, int(origin[0]), int(origin[1]),
                                                          int(origin[2]), int(size), _ptr(out)[0], 0))
        return out

    def blend_block(self, probs, order, size, window, pred, weight, origin):
        """`pred[block] += mean * window; weight[block] += window` (predict.py:244-245) for one block's per-axis
        probabilities (CUDA fp32, layouts of `predict_axis`); `pred` `[D,H,W,C]`, `weight` `[D,H,W]` CUDA fp32."""
        d, h, w = (int(v) for v in weight.shape)
        ptrs = [ctypes.c_void_p(probs[a].data_ptr()) if a in probs and probs[a] is not None else None for a in (0, 1, 2)]
        c_order = (ctypes.c_int * len(order))(*[int(a) for a in order])
        c_org = (ctypes.c_int * 3)(*[int(v) for v in origin])
        g, gmax, lo = window
        gp, _g = _ptr(g)
        with self._lock:
            self._sync_torch(*[p for p in probs.values() if p is not None], pred, weight)
            self._check(self._lib.iu_engine_blend_block(
                self._h, ptrs[0], ptrs[1], ptrs[2], c_order, len(order), int(size), int(self.num_classes), gp, float(gmax),
                float(lo), _ptr(pred)[0], _ptr(weight)[0], d, h, w, c_org, 0))

    # ------------------------------------------------------------------ Zarr staging
    @staticmethod
    def _voxel_layout(shape, chunks, itemsize):
        """(d, h, w, bytes per voxel, bytes per voxel inside a chunk, cz, cy, cx) for an array chunked over its first
        three axes only; one trailing (class) axis may be shorter than the chunk's (padding)."""
        shape, chunks = tuple(int(v) for v in shape), tuple(int(v) for v in chunks)
        ok = len(shape) >= 3 and len(chunks) == len(shape) and \
            (shape[3:] == chunks[3:] or (len(shape) == 4 and 0 < shape[3] <= chunks[3]))
        if not ok:
            raise ValueError(f"chunk layout {chunks} of an array of shape {shape}: only the first three axes may be "
                             f"chunked (the reference chunks (128,128,128,C), predict.py:177)")
        elem = int(np.prod(shape[3:], dtype=np.int64)) * int(itemsize)
        celem = int(np.prod(chunks[3:], dtype=np.int64)) * int(itemsize)
        return shape[:3] + (elem, celem) + chunks[:3]

    def to_chunks(self, volume, chunks, out=None):
        """CUDA array `[D,H,W,...]` -> chunk-major CUDA staging `[n_chunks, *chunks]` (edge padding zeroed), the
        layout `zarr3.Array.write_chunk_major` compresses from."""
        d, h, w, elem, celem, cz, cy, cx = self._voxel_layout(volume.shape, chunks, volume.element_size())
        n = -(-d // cz) * -(-h // cy) * -(-w // cx)
        if out is None:
            out = torch.empty((n,) + tuple(int(v) for v in chunks), dtype=volume.dtype, device=volume.device)
        if out.numel() != n * int(np.prod(chunks, dtype=np.int64)) or out.dtype != volume.dtype:
            raise ValueError("to_chunks: staging tensor has the wrong size / dtype")
        with self._lock:
            self._sync_torch(volume, out)
            self._check(self._lib.iu_engine_to_chunks(self._h, _ptr(volume)[0], d, h, w, elem, celem, cz, cy, cx,
                                                      _ptr(out)[0], 0))
        return out

    def from_chunks(self, staged, shape, chunks, out=None):
        """Chunk-major CUDA staging (as `zarr3.Array.read_chunk_major` decodes it) -> CUDA array of `shape`."""
        d, h, w, elem, celem, cz, cy, cx = self._voxel_layout(shape, chunks, staged.element_size())
        n = -(-d // cz) * -(-h // cy) * -(-w // cx)
        if staged.numel() != n * int(np.prod(chunks, dtype=np.int64)):
            raise ValueError("from_chunks: staging tensor has the wrong size")
        if out is None:
            out = torch.empty(tuple(int(v) for v in shape), dtype=staged.dtype, device=staged.device)
        with self._lock:
            self._sync_torch(staged, out)
            self._check(self._lib.iu_engine_from_chunks(self._h, _ptr(staged)[0], d, h, w, elem, celem, cz, cy, cx,
                                                        _ptr(out)[0], 0))
        return out

    def zoom_nearest(self, src, tables, out=None):
        """One pyramid level on the device: `out[i,j,k(,l)] = src[t0[i], t1[j], t2[k](, t3[l])]`, 0 where a table holds
        -1.  `tables` come from `utils.zoom_tables` (the reference's block-wise `ndimage.zoom(order=0)`)."""
        if src.dim() not in (3, 4) or len(tables) != src.dim():
            raise ValueError("zoom_nearest: 3-D or 4-D arrays with one table per axis")
        tabs = [np.ascontiguousarray(t, dtype=np.int32) for t in tables]
        dshape = tuple(int(t.size) for t in tabs)
        if out is None:
            out = torch.empty(dshape, dtype=src.dtype, device=src.device)
        if tuple(out.shape) != dshape or out.dtype != src.dtype:
            raise ValueError("zoom_nearest: destination has the wrong shape / dtype")
        if out.numel() == 0:
            return out          # an empty level (the reference halves the class axis too: C = 1 -> 0)
        sd = [int(v) for v in src.shape] + [1] * (4 - src.dim())
        dd = list(dshape) + [1] * (4 - src.dim())
        if src.dim() == 3:
            tabs.append(np.zeros(1, np.int32))
        ip = ctypes.POINTER(ctypes.c_int)
        with self._lock:
            self._sync_torch(src, out)
            self._check(self._lib.iu_engine_zoom_nearest(
                self._h, _ptr(src)[0], (ctypes.c_int * 4)(*sd), _ptr(out)[0],
                (ctypes.c_int * 4)(*dd), *[t.ctypes.data_as(ip) for t in tabs], int(src.element_size()), 0))
        return out

    def finalise(self, pred, weight, out_u8=None, out_labels=None):
        """`normalize_shard` (predict.py:252-255) over CUDA fp32 `pred` / `weight` -> CUDA uint8 (and argmax labels)."""
        with self._lock:
            self._sync_torch(pred, weight, out_u8, out_labels)
            self._check(self._lib.iu_engine_finalise(self._h, _ptr(pred)[0], _ptr(weight)[0], int(weight.numel()),
                                                     int(pred.shape[-1]), _ptr(out_u8)[0], _ptr(out_labels)[0], 0))

    # ------------------------------------------------------------------ test hook
    def conv_test(self, src0, src1, weight, bias, ksize, stride, residual=None, relu=True, up2x=False,
                  src0_up=False):
        """One tensor-core conv on 16-bit NHWC CUDA tensors, exactly as the engine runs its layers.
        `src0_up`: src0 is stored at half resolution and read through a 2x nearest upsample (decoder conv1)."""
        b, h, w, c0 = src0.shape
        if src0_up:
            h, w = 2 * h, 2 * w
        c1 = 0 if src1 is None else src1.shape[3]
        cout = weight.shape[0]
        pad = ksize // 2
        oh, ow = (h + 2 * pad - ksize) // stride + 1, (w + 2 * pad - ksize) // stride + 1
        f = 2 if up2x else 1
        if src0.dtype != self.act_dtype:
            raise TypeError(f"engine precision is {self.precision}: tensors must be {self.act_dtype}")
        out = torch.zeros((b, oh * f, ow * f, cout), dtype=self.act_dtype, device=src0.device)
        wt = np.ascontiguousarray(weight, dtype=np.float32)
        bs = np.ascontiguousarray(bias, dtype=np.float32)
        with self._lock:
            self._sync_torch(src0)
            self._check(self._lib.iu_engine_conv_test(
                self._h, _ptr(src0)[0], c0, _ptr(src1)[0], c1, b, h, w, ksize, stride, _ptr(wt)[0], _ptr(bs)[0],
                cout, _ptr(residual)[0], int(relu), int(bool(up2x)) | (2 if src0_up else 0), _ptr(out)[0]))
        return out

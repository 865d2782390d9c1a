"""Zarr v3 stores as the reference reads and writes them, without the `zarr` package.

The reference keeps every volume as a Zarr v3 group of multiscale levels '0', '1', ... whose arrays are created with
`create_array(chunks=(128,)*3, shards=(256,)*3)` (`utils.py:66-71,85-90`, `predict.py:174-198`; format stated at
`README.md:19`, `zarr==3.1.3` pinned at `pyproject.toml:23`).  With zarr-python's defaults that is, on disk:

    <store>/zarr.json                      {"zarr_format": 3, "node_type": "group", "attributes": {}}
    <store>/<level>/zarr.json              array metadata: regular chunk grid whose "chunks" are the SHARDS, default
                                           chunk-key encoding with '/', fill_value 0, one codec `sharding_indexed`
                                           {chunk_shape: inner chunk, codecs: [bytes, zstd(level 0, no checksum)],
                                            index_codecs: [bytes(little), crc32c], index_location: "end"}
    <store>/<level>/c/<i>/<j>/<k>[/<l>]    one file per shard: the encoded inner chunks back to back, then the index
                                           uint64[chunks per shard..., 2] = (offset, nbytes), little endian, C order,
                                           (2^64-1, 2^64-1) for a chunk that equals the fill value, then crc32c(index)

This module implements that subset of the Zarr v3 core + sharding + zstd/gzip/crc32c codec specifications, and the
small part of the zarr-python API the reference touches (`zarr.open(path, mode)`, `group[name]`, `group.array_keys()`,
`group.create_array(name=, shape=, chunks=, shards=, dtype=, overwrite=)`, `array.shape/.chunks/.shards/.dtype`,
basic-slice `__getitem__` / `__setitem__`), so the reference's callers keep working when handed these objects.
`zarr` itself is not installed in the build image, so conformance is pinned by known-answer tests of the pieces the
specifications fix (crc32c, index layout, metadata documents) and by round trips -- see DESIGN.md.

Besides the slicing API there are two bulk paths used by `predict_volumes`, which move only whole inner chunks so that
the byte shuffling happens on the GPU (`iu_engine_to_chunks` / `iu_engine_from_chunks`) and the host threads do nothing
but (de)compress: `Array.read_chunk_major` and `Array.write_chunk_major`.
"""
import ctypes
import ctypes.util
import json
import os
import shutil
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

__all__ = ["open", "open_group", "Group", "Array", "crc32c", "default_pool"]

_builtin_open = open
_EMPTY = 0xFFFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------------------- codecs
def _make_crc32c_table():
    poly = 0x82F63B78  # Castagnoli, reflected
    table = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        table.append(c)
    return table


_CRC32C_TABLE = _make_crc32c_table()


def crc32c(data) -> int:
    """CRC-32C (Castagnoli) of a bytes-like object, as the `crc32c` codec appends it (4 bytes, little endian)."""
    crc = 0xFFFFFFFF
    table = _CRC32C_TABLE
    for b in bytes(data):
        crc = table[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


class _Zstd:
    """libzstd through ctypes (the calls release the GIL, so a thread pool compresses shards in parallel)."""

    def __init__(self):
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        try:
            lib = ctypes.CDLL(name)
        except OSError as e:
            raise RuntimeError("Zarr stores written by the reference are zstd-compressed; libzstd was not found") from e
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        lib.ZSTD_getErrorName.restype = ctypes.c_char_p
        lib.ZSTD_getErrorName.argtypes = [ctypes.c_size_t]
        self.lib = lib

    def compress(self, arr: np.ndarray, level: int) -> bytes:
        """`arr`: C-contiguous array.  Level 0 is zstd's default level, which is what zarr-python writes."""
        n = arr.nbytes
        cap = self.lib.ZSTD_compressBound(n)
        dst = ctypes.create_string_buffer(cap)
        r = self.lib.ZSTD_compress(dst, cap, arr.ctypes.data, n, int(level))
        if self.lib.ZSTD_isError(r):
            raise RuntimeError("zstd: " + self.lib.ZSTD_getErrorName(r).decode())
        return ctypes.string_at(dst, r)

    def decompress_into(self, src, out: np.ndarray):
        """`src`: bytes-like frame; `out`: C-contiguous array that receives exactly `out.nbytes` bytes."""
        buf = np.frombuffer(src, dtype=np.uint8)
        r = self.lib.ZSTD_decompress(out.ctypes.data, out.nbytes, buf.ctypes.data, buf.size)
        if self.lib.ZSTD_isError(r):
            raise RuntimeError("zstd: " + self.lib.ZSTD_getErrorName(r).decode())
        if r != out.nbytes:
            raise RuntimeError(f"zstd: chunk decodes to {r} bytes, {out.nbytes} expected")


_zstd = None


def _get_zstd():
    global _zstd
    if _zstd is None:
        _zstd = _Zstd()
    return _zstd


_pool = None


def default_pool() -> ThreadPoolExecutor:
    """One thread per host core (the reference normalises shards with `Parallel(n_jobs=-1)`, `predict.py:257`)."""
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1), thread_name_prefix="iu-zarr")
    return _pool


_DTYPES = {"bool": "?", "int8": "i1", "int16": "i2", "int32": "i4", "int64": "i8", "uint8": "u1", "uint16": "u2",
           "uint32": "u4", "uint64": "u8", "float16": "f2", "float32": "f4", "float64": "f8"}
_DTYPE_NAMES = {np.dtype(v): k for k, v in _DTYPES.items()}


def _parse_fill(v, dtype):
    if isinstance(v, str):
        v = {"NaN": np.nan, "Infinity": np.inf, "-Infinity": -np.inf}.get(v, v)
    return np.array(v if v is not None else 0).astype(dtype)[()]


class _ByteCodecs:
    """The array->bytes codec (`bytes`) followed by bytes->bytes codecs, for one chunk shape."""

    def __init__(self, codecs, dtype, what):
        names = [c["name"] for c in codecs]
        if not names or names[0] != "bytes":
            raise NotImplementedError(f"{what}: codec chain {names} is not supported (expected 'bytes' first; "
                                      f"'transpose' and other array codecs are not implemented)")
        endian = (codecs[0].get("configuration") or {}).get("endian", "little")
        self.dtype = np.dtype(dtype).newbyteorder("<" if endian == "little" else ">") if np.dtype(dtype).itemsize > 1 \
            else np.dtype(dtype)
        self.steps = []
        for c in codecs[1:]:
            cfg = c.get("configuration") or {}
            if c["name"] == "zstd":
                self.steps.append(("zstd", int(cfg.get("level", 0))))
            elif c["name"] == "gzip":
                self.steps.append(("gzip", int(cfg.get("level", 5))))
            elif c["name"] == "crc32c":
                self.steps.append(("crc32c", 0))
            else:
                raise NotImplementedError(f"{what}: codec '{c['name']}' is not supported (zstd, gzip, crc32c are)")

    def encode(self, arr: np.ndarray) -> bytes:
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        if not self.steps:
            return arr.tobytes()
        data = arr
        for kind, level in self.steps:
            if kind == "zstd":
                data = _get_zstd().compress(data if isinstance(data, np.ndarray) else np.frombuffer(data, np.uint8), level)
            elif kind == "gzip":
                data = zlib.compress(bytes(data) if not isinstance(data, np.ndarray) else data.tobytes(), level, wbits=31)
            else:
                raw = data.tobytes() if isinstance(data, np.ndarray) else bytes(data)
                data = raw + crc32c(raw).to_bytes(4, "little")
        return data

    def decode_into(self, data, out: np.ndarray):
        """Decode one chunk into `out` (C-contiguous, this codec's dtype up to byte order, the chunk's shape)."""
        steps = self.steps[::-1]
        for n, (kind, _) in enumerate(steps):
            last = n == len(steps) - 1
            if kind == "crc32c":
                body, tail = bytes(data[:-4]), bytes(data[-4:])
                if crc32c(body) != int.from_bytes(tail, "little"):
                    raise RuntimeError("crc32c mismatch in a chunk")
                data = body
            elif kind == "gzip":
                data = zlib.decompress(bytes(data), wbits=47)
            elif last and self.dtype.isnative or last and self.dtype.itemsize == 1:
                _get_zstd().decompress_into(data, out)
                return
            else:
                tmp = np.empty(out.nbytes, np.uint8)
                _get_zstd().decompress_into(data, tmp)
                data = tmp
        flat = np.frombuffer(data, dtype=self.dtype, count=out.size)
        out[...] = flat.reshape(out.shape)


# ------------------------------------------------------------------------------------------------- arrays
class Array:
    """One Zarr v3 array directory.  `chunks` / `shards` follow zarr-python's naming: with the sharding codec `shards`
    is the shape of one stored file and `chunks` the shape of the independently compressed pieces inside it; without
    it `shards` is None and `chunks` is the stored chunk."""

    def __init__(self, path, meta, writable):
        self.path = path
        self.meta = meta
        self.writable = writable
        if meta.get("zarr_format") != 3 or meta.get("node_type") != "array":
            raise ValueError(f"{path}: not a Zarr v3 array")
        self.shape = tuple(int(v) for v in meta["shape"])
        self.ndim = len(self.shape)
        if meta["data_type"] not in _DTYPES:
            raise NotImplementedError(f"{path}: data type {meta['data_type']!r} is not supported")
        self.dtype = np.dtype(_DTYPES[meta["data_type"]])
        grid = meta["chunk_grid"]
        if grid["name"] != "regular":
            raise NotImplementedError(f"{path}: chunk grid {grid['name']!r}")
        self._outer = tuple(int(v) for v in grid["configuration"]["chunk_shape"])
        enc = meta.get("chunk_key_encoding", {"name": "default"})
        self._key_default = enc["name"] == "default"
        self._sep = (enc.get("configuration") or {}).get("separator", "/" if self._key_default else ".")
        self.fill_value = _parse_fill(meta.get("fill_value", 0), self.dtype)
        if meta.get("storage_transformers"):
            raise NotImplementedError(f"{path}: storage transformers")
        codecs = meta["codecs"]
        if len(codecs) == 1 and codecs[0]["name"] == "sharding_indexed":
            cfg = codecs[0]["configuration"]
            self.shards = self._outer
            self.chunks = tuple(int(v) for v in cfg["chunk_shape"])
            if any(s % c for s, c in zip(self.shards, self.chunks)):
                raise ValueError(f"{path}: shard shape {self.shards} is not a multiple of chunk shape {self.chunks}")
            self._inner = _ByteCodecs(cfg["codecs"], self.dtype, path)
            self._index = _ByteCodecs(cfg.get("index_codecs", [{"name": "bytes"}, {"name": "crc32c"}]), np.uint64, path)
            self._index_at_end = cfg.get("index_location", "end") == "end"
        else:
            self.shards = None
            self.chunks = self._outer
            self._inner = _ByteCodecs(codecs, self.dtype, path)
            self._index = None
        self._per_shard = tuple(s // c for s, c in zip(self._outer, self.chunks))      # inner chunks per stored file
        self.chunk_grid = tuple(-(-n // c) for n, c in zip(self.shape, self.chunks))   # inner chunks over the array
        self.shard_grid = tuple(-(-n // s) for n, s in zip(self.shape, self._outer))

    # ---- zarr-python-like surface
    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"<zarr3.Array {self.path} shape={self.shape} chunks={self.chunks} shards={self.shards} dtype={self.dtype}>"

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a.astype(dtype) if dtype is not None else a

    # ---- stored files
    def _file(self, coords):
        parts = [str(int(c)) for c in coords]
        if self._key_default:
            return os.path.join(self.path, "c", *parts) if self._sep == "/" else \
                os.path.join(self.path, self._sep.join(["c"] + parts))
        return os.path.join(self.path, *parts) if self._sep == "/" else os.path.join(self.path, self._sep.join(parts))

    def _index_nbytes(self):
        n = int(np.prod(self._per_shard)) * 16
        return n + 4 * sum(1 for k, _ in self._index.steps if k == "crc32c")

    def _read_file(self, coords):
        try:
            with _builtin_open(self._file(coords), "rb") as f:
                return f.read()
        except FileNotFoundError:
            return None

    def _split_shard(self, blob):
        """Stored file -> (index uint64[n_inner, 2], memoryview of the file)."""
        n = self._index_nbytes()
        if len(blob) < n:
            raise RuntimeError(f"{self.path}: shard file shorter than its index")
        raw = blob[-n:] if self._index_at_end else blob[:n]
        index = np.empty((int(np.prod(self._per_shard)), 2), np.uint64)
        self._index.decode_into(raw, index)
        return index, memoryview(blob)

    def _decode_inner(self, blob_view, index, k, out):
        off, nb = int(index[k, 0]), int(index[k, 1])
        if off == _EMPTY and nb == _EMPTY:
            out[...] = self.fill_value
        else:
            self._inner.decode_into(blob_view[off:off + nb], out)

    def read_stored(self, coords, out=None):
        """Decode the stored chunk / shard at grid `coords` into an array of shape `shards or chunks`."""
        if out is None:
            out = np.empty(self._outer, self.dtype)
        blob = self._read_file(coords)
        if blob is None:
            out[...] = self.fill_value
            return out
        if self.shards is None:
            tmp = out if out.flags.c_contiguous else np.empty(self._outer, self.dtype)
            self._inner.decode_into(blob, tmp)
            if tmp is not out:
                out[...] = tmp
            return out
        index, view = self._split_shard(blob)
        tmp = np.empty(self.chunks, self.dtype)
        for k, ic in enumerate(np.ndindex(*self._per_shard)):
            self._decode_inner(view, index, k, tmp)
            out[tuple(slice(i * c, (i + 1) * c) for i, c in zip(ic, self.chunks))] = tmp
        return out

    def _encode_shard(self, pieces):
        """`pieces`: inner chunks (C-contiguous arrays, or None for "equals the fill value") in C order of the shard's
        inner grid -> file bytes, or None when every piece is empty (zarr-python then deletes / skips the file)."""
        return self._assemble_shard([None if p is None else self._inner.encode(p) for p in pieces])

    def _assemble_shard(self, encoded):
        """Encoded inner chunks (bytes, or None for empty) in the shard's C order -> file bytes (or None)."""
        n = len(encoded)
        index = np.full((n, 2), _EMPTY, np.uint64)
        body = []
        pos = 0 if self._index_at_end else self._index_nbytes()
        for k, enc in enumerate(encoded):
            if enc is None:
                continue
            index[k] = (pos, len(enc))
            body.append(enc)
            pos += len(enc)
        if not body:
            return None
        idx = self._index.encode(index)
        return b"".join(body + [idx]) if self._index_at_end else b"".join([idx] + body)

    def _is_fill(self, a):
        if isinstance(self.fill_value, np.floating) and np.isnan(self.fill_value):
            return bool(np.isnan(a).all())
        if a.dtype.itemsize == 1 and self.fill_value == 0:
            return not a.any()
        return bool((a == self.fill_value).all())

    def write_stored(self, coords, data):
        """Encode one whole stored chunk / shard (`data` of shape `shards or chunks`, edge padding included)."""
        if not self.writable:
            raise PermissionError(f"{self.path} was opened read-only")
        data = np.asarray(data)
        if self.shards is None:
            blob = None if self._is_fill(data) else self._inner.encode(data)
        else:
            pieces = []
            for ic in np.ndindex(*self._per_shard):
                p = np.ascontiguousarray(data[tuple(slice(i * c, (i + 1) * c) for i, c in zip(ic, self.chunks))])
                pieces.append(None if self._is_fill(p) else p)
            blob = self._encode_shard(pieces)
        self._store_file(coords, blob)

    def _store_file(self, coords, blob):
        fn = self._file(coords)
        if blob is None:
            try:
                os.remove(fn)
            except FileNotFoundError:
                pass
            return
        os.makedirs(os.path.dirname(fn), exist_ok=True)
        tmp = fn + f".partial.{os.getpid()}"
        with _builtin_open(tmp, "wb") as f:
            f.write(blob)
        os.replace(tmp, fn)

    # ---- basic slicing (what `get_padded_block` and `resize_volume` do to a zarr array)
    def _normalise_key(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            i = [n for n, k in enumerate(key) if k is Ellipsis][0]
            key = key[:i] + (slice(None),) * (self.ndim - (len(key) - 1)) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        if len(key) != self.ndim:
            raise IndexError(f"too many indices for array: array is {self.ndim}-dimensional")
        sel, squeeze = [], []
        for k, n in zip(key, self.shape):
            if isinstance(k, slice):
                start, stop, step = k.indices(n)
                if step != 1:
                    raise NotImplementedError("zarr3.Array: only unit-step slices are supported")
                sel.append((start, max(start, stop)))
            else:
                i = int(k)
                if i < -n or i >= n:
                    raise IndexError(f"index {i} is out of bounds for axis with size {n}")
                i %= n
                sel.append((i, i + 1))
                squeeze.append(len(sel) - 1)
        return sel, tuple(squeeze)

    def _touched(self, sel):
        ranges = [range(a // s, -(-b // s)) if b > a else range(0) for (a, b), s in zip(sel, self._outer)]
        return np.ndindex(*[len(r) for r in ranges]), ranges

    def __getitem__(self, key):
        sel, squeeze = self._normalise_key(key)
        out = np.empty([b - a for a, b in sel], self.dtype)
        if out.size:
            it, ranges = self._touched(sel)
            for rel in it:
                sc = tuple(r[i] for r, i in zip(ranges, rel))
                stored = self.read_stored(sc)
                src, dst = [], []
                for (a, b), s, c in zip(sel, self._outer, sc):
                    lo, hi = max(a, c * s), min(b, (c + 1) * s)
                    src.append(slice(lo - c * s, hi - c * s))
                    dst.append(slice(lo - a, hi - a))
                out[tuple(dst)] = stored[tuple(src)]
        return out.squeeze(axis=squeeze) if squeeze else out

    def __setitem__(self, key, value):
        sel, squeeze = self._normalise_key(key)
        shape = [b - a for a, b in sel]
        value = np.asarray(value)
        if squeeze:
            value = np.expand_dims(value, squeeze) if value.ndim == len(shape) - len(squeeze) else value
        value = np.broadcast_to(value.astype(self.dtype, copy=False), shape)     # numpy's own error text on mismatch
        if not value.size:
            return
        it, ranges = self._touched(sel)
        for rel in it:
            sc = tuple(r[i] for r, i in zip(ranges, rel))
            src, dst, whole = [], [], True
            for (a, b), s, c, n in zip(sel, self._outer, sc, self.shape):
                lo, hi = max(a, c * s), min(b, (c + 1) * s)
                whole &= lo == c * s and hi == min((c + 1) * s, n)
                dst.append(slice(lo - c * s, hi - c * s))
                src.append(slice(lo - a, hi - a))
            if whole:
                stored = np.full(self._outer, self.fill_value, self.dtype)
            else:
                stored = self.read_stored(sc)
            stored[tuple(dst)] = value[tuple(src)]
            self.write_stored(sc, stored)

    # ---- bulk paths: whole inner chunks only
    def chunk_major_shape(self):
        """Shape of the chunk-major staging buffer: (number of inner chunks over the array, *chunk shape)."""
        return (int(np.prod(self.chunk_grid)),) + self.chunks

    def _shard_chunk_ids(self, sc):
        """Inner chunks of stored file `sc` in the file's C order -> flat ids in the array-wide inner chunk grid, or
        -1 for inner chunks that lie wholly outside the array."""
        ids = []
        for ic in np.ndindex(*self._per_shard):
            g = tuple(s * p + i for s, p, i in zip(sc, self._per_shard, ic))
            ids.append(int(np.ravel_multi_index(g, self.chunk_grid)) if all(a < b for a, b in zip(g, self.chunk_grid))
                       else -1)
        return ids

    def read_chunk_major(self, out=None, pool=None):
        """Decode the whole array into chunk-major order: `out[id]` is inner chunk `id` (C order over `chunk_grid`),
        edge chunks padded with whatever was stored (the fill value).  Only decompression happens on the host -- one
        pool task per stored file to read it and parse its index, then one per inner chunk to decompress; the scatter
        into `[D,H,W,...]` order is `Engine.from_chunks` on the device."""
        if out is None:
            out = np.empty(self.chunk_major_shape(), self.dtype)
        if tuple(out.shape) != self.chunk_major_shape() or out.dtype != self.dtype or not out.flags.c_contiguous:
            raise ValueError("read_chunk_major: staging buffer has the wrong shape / dtype / layout")
        pool = pool or default_pool()

        def open_file(sc):
            ids = self._shard_chunk_ids(sc)
            blob = self._read_file(sc)
            if blob is None:
                return [(None, None, 0, i) for i in ids if i >= 0]
            if self.shards is None:
                return [(blob, None, 0, ids[0])]
            index, view = self._split_shard(blob)
            return [(view, index, k, i) for k, i in enumerate(ids) if i >= 0]

        def decode(job):
            view, index, k, i = job
            if view is None:
                out[i] = self.fill_value
            elif index is None:
                self._inner.decode_into(view, out[i])
            else:
                self._decode_inner(view, index, k, out[i])

        jobs = [j for js in pool.map(open_file, list(np.ndindex(*self.shard_grid))) for j in js]
        list(pool.map(decode, jobs))
        return out

    def write_chunk_major(self, staged, pool=None, wait=True):
        """Inverse of `read_chunk_major`: `staged[id]` holds inner chunk `id` with its edge padding already set to the
        fill value (`Engine.to_chunks` writes zeros there).  One pool task per inner chunk (test for all-fill,
        compress), then one per stored file (index, checksum, write).  With `wait=False` the list of futures of the
        file tasks is returned instead of being waited for; `staged` must stay untouched until they are done."""
        if not self.writable:
            raise PermissionError(f"{self.path} was opened read-only")
        if tuple(staged.shape) != self.chunk_major_shape() or staged.dtype != self.dtype:
            raise ValueError("write_chunk_major: staging buffer has the wrong shape / dtype")
        pool = pool or default_pool()

        def encode(i):
            p = staged[i]
            return None if self._is_fill(p) else self._inner.encode(p)

        encoded = [pool.submit(encode, i) for i in range(staged.shape[0])]

        def store(sc):
            ids = self._shard_chunk_ids(sc)
            pieces = [None if i < 0 else encoded[i].result() for i in ids]
            if self.shards is None:
                self._store_file(sc, pieces[0])
            else:
                self._store_file(sc, self._assemble_shard(pieces))

        # file tasks wait on chunk tasks submitted before them, so a FIFO pool cannot deadlock
        files = [pool.submit(store, sc) for sc in np.ndindex(*self.shard_grid)]
        if not wait:
            return files
        for f in files:
            f.result()
        return None


# ------------------------------------------------------------------------------------------------- groups
def _array_metadata(shape, chunks, shards, dtype, fill_value=0):
    dtype = np.dtype(dtype)
    if dtype not in _DTYPE_NAMES:
        raise NotImplementedError(f"dtype {dtype} is not supported")
    one_byte = dtype.itemsize == 1
    inner = [{"name": "bytes"} if one_byte else {"name": "bytes", "configuration": {"endian": "little"}},
             {"name": "zstd", "configuration": {"level": 0, "checksum": False}}]
    if shards is None:
        outer, codecs = chunks, inner
    else:
        outer = shards
        codecs = [{"name": "sharding_indexed", "configuration": {
            "chunk_shape": [int(v) for v in chunks], "codecs": inner,
            "index_codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "crc32c"}],
            "index_location": "end"}}]
    fv = fill_value.item() if isinstance(fill_value, np.generic) else fill_value
    if dtype.kind == "f":
        fv = float(fv)
    elif dtype.kind == "b":
        fv = bool(fv)
    else:
        fv = int(fv)
    return {"shape": [int(v) for v in shape], "data_type": _DTYPE_NAMES[dtype],
            "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [int(v) for v in outer]}},
            "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
            "fill_value": fv, "codecs": codecs, "attributes": {}, "zarr_format": 3, "node_type": "array",
            "storage_transformers": []}


def _read_json(path):
    with _builtin_open(path, "r") as f:
        return json.load(f)


def _write_json(path, doc):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with _builtin_open(path, "w") as f:
        json.dump(doc, f, indent=2)


class Group:
    def __init__(self, path, writable):
        self.path = path
        self.writable = writable

    def __repr__(self):
        return f"<zarr3.Group {self.path}>"

    def _members(self, node_type):
        out = []
        if os.path.isdir(self.path):
            for name in sorted(os.listdir(self.path)):
                mj = os.path.join(self.path, name, "zarr.json")
                if os.path.isfile(mj):
                    try:
                        if _read_json(mj).get("node_type") == node_type:
                            out.append(name)
                    except (OSError, ValueError):
                        pass
        return out

    def array_keys(self):
        return iter(self._members("array"))

    def group_keys(self):
        return iter(self._members("group"))

    def __contains__(self, name):
        return os.path.isfile(os.path.join(self.path, str(name), "zarr.json"))

    def __getitem__(self, name):
        p = os.path.join(self.path, str(name))
        mj = os.path.join(p, "zarr.json")
        if not os.path.isfile(mj):
            raise KeyError(name)
        meta = _read_json(mj)
        return Group(p, self.writable) if meta.get("node_type") == "group" else Array(p, meta, self.writable)

    def create_array(self, name, shape, chunks, shards=None, dtype="uint8", overwrite=False, fill_value=0, **unused):
        """`Group.create_array` with zarr-python's defaults for everything the reference leaves unset (`bytes` + zstd
        level 0 inside a `sharding_indexed` codec when `shards` is given)."""
        if not self.writable:
            raise PermissionError(f"{self.path} was opened read-only")
        shape, chunks = tuple(int(v) for v in shape), tuple(int(v) for v in chunks)
        if len(chunks) != len(shape) or (shards is not None and len(shards) != len(shape)):
            raise ValueError("create_array: chunks / shards must have one entry per dimension")
        if shards is not None:
            shards = tuple(int(v) for v in shards)
            if any(s % c for s, c in zip(shards, chunks)):
                raise ValueError(f"The array's `chunk_shape` {shards} needs to be divisible by the shard's inner "
                                 f"`chunk_shape` {chunks}.")
        p = os.path.join(self.path, str(name))
        if os.path.exists(p):
            if not overwrite:
                raise FileExistsError(f"{p} exists (pass overwrite=True)")
            shutil.rmtree(p)
        meta = _array_metadata(shape, chunks, shards, dtype, fill_value)
        _write_json(os.path.join(p, "zarr.json"), meta)
        return Array(p, meta, True)


def open_group(path, mode="r"):
    return open(path, mode=mode)


def open(path, mode="r"):
    """`zarr.open(path, mode)` for local directory stores: 'r' read-only, 'r+' read/write existing, 'a' read/write or
    create, 'w' create (replacing whatever is at `path`) -- `utils.py:21,53,79`, `predict.py:167,172,181,191`."""
    path = os.fspath(path)
    mj = os.path.join(path, "zarr.json")
    if mode == "w":
        if os.path.isdir(path):
            shutil.rmtree(path)
        elif os.path.exists(path):
            os.remove(path)
        _write_json(mj, {"attributes": {}, "zarr_format": 3, "node_type": "group"})
        return Group(path, True)
    if mode not in ("r", "r+", "a"):
        raise ValueError(f"zarr3.open: unsupported mode {mode!r}")
    if not os.path.isfile(mj):
        if mode == "a":
            _write_json(mj, {"attributes": {}, "zarr_format": 3, "node_type": "group"})
            return Group(path, True)
        if os.path.isfile(os.path.join(path, ".zgroup")) or os.path.isfile(os.path.join(path, ".zarray")):
            raise NotImplementedError(f"{path} is a Zarr v2 store; the reference writes Zarr v3 (README.md:19)")
        raise FileNotFoundError(f"{path}: no Zarr v3 node here (zarr.json missing)")
    meta = _read_json(mj)
    if meta.get("zarr_format") != 3:
        raise NotImplementedError(f"{path}: zarr_format {meta.get('zarr_format')}")
    writable = mode != "r"
    return Array(path, meta, writable) if meta.get("node_type") == "array" else Group(path, writable)

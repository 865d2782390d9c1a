"""Zarr v3 store I/O and the pyramid's host logic -- no GPU.

The `zarr` package is absent from the image, so the on-disk format is pinned by what the specifications fix: CRC-32C
known answers (RFC 3720 B.4), the shard index layout, the metadata documents, and a second zstd binding (pyarrow)
decoding the frames this code writes.  The pyramid's index tables are pinned by `tests/golden/multiscales.npz`,
recorded from the verbatim reference `utils.add_multiscales` (`oracle/make_golden.py`)."""
import json
import os
import struct

import numpy as np
import pytest

import interactive_unet_b200  # noqa: F401  (registers the package alias)
from interactive_unet_b200 import utils as iu_utils
from interactive_unet_b200 import zarr3
from oracle import predict_port

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MS_CASES = ["c2", "c4", "ragged_c2", "image", "fill_c2", "u16", "odd", "c3", "small"]


def test_crc32c_known_answers():
    assert zarr3.crc32c(b"123456789") == 0xE3069283
    assert zarr3.crc32c(bytes(32)) == 0x8A9136AA                    # RFC 3720 B.4
    assert zarr3.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert zarr3.crc32c(bytes(range(32))) == 0x46DD794E
    assert zarr3.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert zarr3.crc32c(b"") == 0


def test_metadata_documents(tmp_path):
    root = zarr3.open(tmp_path / "v.zarr", mode="w")
    root.create_array(name="0", shape=[40, 36, 44, 2], chunks=(16, 16, 16, 2), shards=(32, 32, 32, 2), dtype="uint8",
                      overwrite=True)
    root.create_array(name="f", shape=(8, 8), chunks=(4, 4), dtype="float32")
    g = json.load(open(tmp_path / "v.zarr" / "zarr.json"))
    assert g == {"attributes": {}, "zarr_format": 3, "node_type": "group"}
    a = json.load(open(tmp_path / "v.zarr" / "0" / "zarr.json"))
    assert a == {
        "shape": [40, 36, 44, 2], "data_type": "uint8",
        "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [32, 32, 32, 2]}},
        "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
        "fill_value": 0,
        "codecs": [{"name": "sharding_indexed", "configuration": {
            "chunk_shape": [16, 16, 16, 2],
            "codecs": [{"name": "bytes"}, {"name": "zstd", "configuration": {"level": 0, "checksum": False}}],
            "index_codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "crc32c"}],
            "index_location": "end"}}],
        "attributes": {}, "zarr_format": 3, "node_type": "array", "storage_transformers": []}
    f = json.load(open(tmp_path / "v.zarr" / "f" / "zarr.json"))
    assert f["codecs"] == [{"name": "bytes", "configuration": {"endian": "little"}},
                           {"name": "zstd", "configuration": {"level": 0, "checksum": False}}]
    assert f["fill_value"] == 0.0 and f["data_type"] == "float32"
    arr = zarr3.open(tmp_path / "v.zarr", mode="r")["0"]
    assert arr.shape == (40, 36, 44, 2) and arr.chunks == (16, 16, 16, 2) and arr.shards == (32, 32, 32, 2)
    assert arr.dtype == np.uint8 and sorted(zarr3.open(tmp_path / "v.zarr").array_keys()) == ["0", "f"]


def test_shard_file_layout_by_hand(tmp_path):
    """Parse a shard file with nothing but `struct`, crc32c and pyarrow's zstd: index at the end, C order,
    (offset, nbytes) uint64 LE, all-ones for chunks equal to the fill value, crc32c of the index last."""
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(0)
    vol = rng.integers(0, 4, (20, 16, 24), dtype=np.uint8)
    vol[:8, :8, 8:16] = 0                                             # inner chunk (0,0,1) of shard (0,0,0) is all fill
    root = zarr3.open(tmp_path / "s.zarr", mode="w")
    arr = root.create_array(name="0", shape=vol.shape, chunks=(8, 8, 8), shards=(16, 16, 16), dtype="uint8")
    arr[:] = vol
    files = sorted(os.path.relpath(os.path.join(d, f), arr.path) for d, _, fs in os.walk(arr.path) for f in fs)
    assert files == ["c/0/0/0", "c/0/0/1", "c/1/0/0", "c/1/0/1", "zarr.json"]
    blob = open(os.path.join(arr.path, "c", "0", "0", "0"), "rb").read()
    index_bytes, crc = blob[-(8 * 16 + 4):-4], blob[-4:]
    assert zarr3.crc32c(index_bytes) == struct.unpack("<I", crc)[0]
    index = np.frombuffer(index_bytes, "<u8").reshape(2, 2, 2, 2)
    assert tuple(index[0, 0, 1]) == (2 ** 64 - 1, 2 ** 64 - 1)
    pos = 0
    for ic in np.ndindex(2, 2, 2):
        if ic == (0, 0, 1):
            continue
        off, nb = (int(v) for v in index[ic])
        assert off == pos                                             # chunks are packed back to back from offset 0
        pos += nb
        raw = pa.decompress(blob[off:off + nb], decompressed_size=512, codec="zstd").to_pybytes()
        expect = vol[tuple(slice(8 * i, 8 * i + 8) for i in ic)]
        assert np.array_equal(np.frombuffer(raw, np.uint8).reshape(8, 8, 8), expect)
    assert pos == len(blob) - (8 * 16 + 4)
    # shard (1,0,1) covers z 16..20 only: its lower inner chunks are padding and must be marked empty
    blob = open(os.path.join(arr.path, "c", "1", "0", "1"), "rb").read()
    index = np.frombuffer(blob[-(8 * 16 + 4):-4], "<u8").reshape(2, 2, 2, 2)
    assert (index[1] == 2 ** 64 - 1).all() and (index[0, :, 1] == 2 ** 64 - 1).all()
    # and frames written by the other binding are readable here
    chunk = np.ascontiguousarray(vol[:8, :8, :8])
    enc = pa.compress(chunk.tobytes(), codec="zstd", asbytes=True)
    out = np.empty((8, 8, 8), np.uint8)
    zarr3._get_zstd().decompress_into(enc, out)
    assert np.array_equal(out, chunk)


@pytest.mark.parametrize("shape,chunks,shards,dtype", [
    ((40, 36, 44), (16, 16, 16), (32, 32, 32), "uint8"),
    ((40, 36, 44, 2), (16, 16, 16, 2), (32, 32, 32, 2), "uint8"),
    ((33, 17, 9, 3), (8, 8, 8, 3), (16, 16, 8, 3), "float32"),
    ((20, 20, 20), (8, 8, 8), None, "uint16"),
    ((5, 7), (4, 4), (8, 8), "int32"),
])
def test_round_trip_and_slicing(tmp_path, shape, chunks, shards, dtype):
    rng = np.random.default_rng(1)
    data = (rng.random(shape) * 200).astype(dtype)
    root = zarr3.open(tmp_path / "r.zarr", mode="w")
    arr = root.create_array(name="0", shape=shape, chunks=chunks, shards=shards, dtype=dtype, overwrite=True)
    assert np.array_equal(arr[...], np.zeros(shape, dtype))          # nothing stored yet: fill value
    arr[:] = data
    again = zarr3.open(tmp_path / "r.zarr", mode="r")["0"]
    assert np.array_equal(again[...], data) and np.array_equal(np.asarray(again), data)
    key = tuple(slice(1, max(2, n - 2)) for n in shape)
    assert np.array_equal(again[key], data[key])
    assert np.array_equal(again[3], data[3]) and np.array_equal(again[-1, 2:4], data[-1, 2:4])
    # read-modify-write of a region that straddles stored files, as `pred[i0:i1, ...] += ...` does (predict.py:244)
    patch = (rng.random([b.stop - b.start for b in key]) * 50).astype(dtype)
    arr[key] = arr[key] + patch
    data[key] = data[key] + patch
    assert np.array_equal(zarr3.open(tmp_path / "r.zarr")["0"][...], data)
    with pytest.raises(PermissionError):
        again[...] = 0
    with pytest.raises(ValueError):
        arr[key] = np.zeros([3] * len(shape), dtype)                 # numpy's broadcast error


def test_chunk_major_bulk_paths(tmp_path):
    rng = np.random.default_rng(2)
    data = rng.integers(0, 256, (40, 36, 44, 2), dtype=np.uint8)
    data[:16, :16, :16] = 0
    root = zarr3.open(tmp_path / "b.zarr", mode="w")
    arr = root.create_array(name="0", shape=data.shape, chunks=(16, 16, 16, 2), shards=(32, 32, 32, 2), dtype="uint8")
    assert arr.chunk_grid == (3, 3, 3, 1) and arr.chunk_major_shape() == (27, 16, 16, 16, 2)
    staged = np.zeros(arr.chunk_major_shape(), np.uint8)
    for n, (gz, gy, gx) in enumerate(np.ndindex(3, 3, 3)):
        piece = data[gz * 16:(gz + 1) * 16, gy * 16:(gy + 1) * 16, gx * 16:(gx + 1) * 16]
        staged[n, :piece.shape[0], :piece.shape[1], :piece.shape[2]] = piece
    arr.write_chunk_major(staged)
    assert np.array_equal(zarr3.open(tmp_path / "b.zarr")["0"][...], data)         # readable through the slicing path
    assert np.array_equal(arr.read_chunk_major(), staged)
    other = root.create_array(name="1", shape=data.shape, chunks=(16, 16, 16, 2), shards=(32, 32, 32, 2), dtype="uint8")
    other[:] = data                                                                 # written through the slicing path
    assert np.array_equal(other.read_chunk_major(), staged)
    for sub in ("0", "1"):                                                          # identical files either way
        a = open(os.path.join(tmp_path, "b.zarr", "0", "c", "1", "1", "1", "0"), "rb").read()
        b = open(os.path.join(tmp_path, "b.zarr", sub, "c", "1", "1", "1", "0"), "rb").read()
        assert a == b


def test_reads_other_codec_choices(tmp_path):
    """Stores written with other settings of the same specification: index at the start, gzip, no compression,
    explicit endian on a one-byte type, '.'-separated and v2-style chunk keys."""
    rng = np.random.default_rng(3)
    data = rng.integers(0, 1000, (12, 10), dtype=np.uint16)
    for n, (inner, loc, enc) in enumerate([
            ([{"name": "bytes", "configuration": {"endian": "big"}}, {"name": "gzip", "configuration": {"level": 1}}],
             "start", {"name": "default", "configuration": {"separator": "."}}),
            ([{"name": "bytes", "configuration": {"endian": "little"}}], "end",
             {"name": "v2", "configuration": {"separator": "."}}),
            ([{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "zstd", "configuration": {"level": 3}},
              {"name": "crc32c"}], "end", {"name": "v2", "configuration": {"separator": "/"}})]):
        root = zarr3.open(tmp_path / f"o{n}.zarr", mode="w")
        arr = root.create_array(name="0", shape=data.shape, chunks=(4, 4), shards=(8, 8), dtype="uint16")
        meta = arr.meta
        meta["codecs"][0]["configuration"].update(codecs=inner, index_location=loc)
        meta["chunk_key_encoding"] = enc
        json.dump(meta, open(os.path.join(arr.path, "zarr.json"), "w"))
        arr = zarr3.open(tmp_path / f"o{n}.zarr", mode="r+")["0"]
        arr[...] = data
        assert np.array_equal(zarr3.open(tmp_path / f"o{n}.zarr")["0"][...], data)
    names = sorted(os.listdir(tmp_path / "o0.zarr" / "0"))
    assert names == ["c.0.0", "c.0.1", "c.1.0", "c.1.1", "zarr.json"]
    assert sorted(os.listdir(tmp_path / "o1.zarr" / "0")) == ["0.0", "0.1", "1.0", "1.1", "zarr.json"]
    meta["codecs"][0]["configuration"]["codecs"] = [{"name": "transpose", "configuration": {"order": [1, 0]}},
                                                    {"name": "bytes"}]
    json.dump(meta, open(os.path.join(arr.path, "zarr.json"), "w"))
    with pytest.raises(NotImplementedError):
        zarr3.open(tmp_path / "o2.zarr")["0"]
    with pytest.raises(FileNotFoundError):
        zarr3.open(tmp_path / "missing.zarr")


def test_corrupt_index_is_detected(tmp_path):
    root = zarr3.open(tmp_path / "c.zarr", mode="w")
    arr = root.create_array(name="0", shape=(8, 8, 8), chunks=(4, 4, 4), shards=(8, 8, 8), dtype="uint8")
    arr[:] = np.arange(512, dtype=np.uint8).reshape(8, 8, 8)
    fn = os.path.join(arr.path, "c", "0", "0", "0")
    blob = bytearray(open(fn, "rb").read())
    blob[-10] ^= 1
    open(fn, "wb").write(bytes(blob))
    with pytest.raises(RuntimeError, match="crc32c"):
        arr[...]


def _golden_ms(name):
    z = np.load(os.path.join(GOLDEN, "multiscales.npz"))
    levels = {int(k): z[f"{name}_level{k}"] for k in z[f"{name}_levels"]}
    return z[f"{name}_volume"], [int(v) for v in z[f"{name}_grid"]], levels, tuple(str(v) for v in z[f"{name}_error"])


def _apply_tables(src, tables):
    out = np.zeros([t.size for t in tables], src.dtype)
    if out.size:
        ok = np.ix_(*[t >= 0 for t in tables])
        out[ok] = src[np.ix_(*[t[t >= 0] for t in tables])]
    return out


@pytest.mark.parametrize("name", MS_CASES)
def test_zoom_tables_match_reference_pyramid(name):
    """`utils.zoom_tables` (host half of the device zoom) applied with numpy == the verbatim reference's levels, and
    raises the reference's ValueError (same message) where `add_multiscales` does."""
    vol, (chunk, shard), levels, (etype, emsg) = _golden_ms(name)
    tail = vol.shape[3:]
    steps = iu_utils._num_steps(vol.shape, (chunk,) * 3 + tail, 0.5)
    cur, made = vol, 0
    try:
        for i in range(steps):
            dst_shape = tuple(int(x * 0.5) for x in cur.shape)
            cur = _apply_tables(cur, iu_utils.zoom_tables(cur.shape, dst_shape, 0.5, shard))
            assert cur.shape == levels[i + 1].shape
            assert np.array_equal(cur, levels[i + 1]), f"level {i + 1}"
            made += 1
    except ValueError as e:
        assert etype == "ValueError" and str(e) == emsg
        assert made == len(levels) - 1                  # the reference had created (not filled) the failing level
    else:
        assert etype in ("", "UnboundLocalError")       # zero steps: the reference trips over `del z0` (utils.py:77)
        assert made == len(levels)
    if name == "fill_c2":
        assert (levels[1][15] == 0).all() and (levels[1][14] != 0).any()      # scipy's constant-fill plane is real


@pytest.mark.parametrize("name", MS_CASES)
def test_oracle_pyramid_port_matches_golden(name):
    vol, (chunk, shard), levels, (etype, emsg) = _golden_ms(name)
    tail = vol.shape[3:]
    try:
        got = predict_port.multiscale_levels(vol, (chunk,) * 3 + tail, (shard,) * 3 + tail)
    except ValueError as e:
        assert etype == "ValueError" and str(e) == emsg
    else:
        assert etype != "ValueError" and len(got) == len(levels)
        for i, g in enumerate(got):
            assert np.array_equal(g, levels[i + 1])


def test_zoom_axis_matches_scipy_everywhere():
    ndimage = pytest.importorskip("scipy.ndimage")
    for n in list(range(1, 300)) + [384, 512, 640, 800, 1024, 2048]:
        want = ndimage.zoom(np.arange(1, n + 1, dtype=np.int64), 0.5, order=0)
        idx = iu_utils._zoom_axis(n, 0.5)
        assert np.array_equal(np.where(idx >= 0, idx + 1, 0), want), n


def test_read_volume_clips_level(tmp_path):
    root = zarr3.open(tmp_path / "p.zarr", mode="w")
    for lvl, n in enumerate((16, 8, 4)):
        root.create_array(name=str(lvl), shape=(n, n, n), chunks=(4, 4, 4), shards=(8, 8, 8), dtype="uint8")
    assert iu_utils.read_volume(tmp_path / "p.zarr", level=1).shape == (8, 8, 8)
    assert iu_utils.read_volume(tmp_path / "p.zarr", level=-3).shape == (16, 16, 16)
    with pytest.raises(KeyError):                       # utils.py:25 clips to num_scales, one past the last level
        iu_utils.read_volume(tmp_path / "p.zarr", level=9)


def test_random_stores_round_trip_property(tmp_path_factory):
    """Property test (hypothesis): any shape / chunking / shard multiple / region write sequence reads back as the
    numpy model of the same assignments, through both the slicing and the chunk-major paths."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @st.composite
    def stores(draw):
        ndim = draw(st.integers(1, 4))
        chunks = tuple(draw(st.integers(1, 5)) for _ in range(ndim))
        mult = tuple(draw(st.integers(1, 3)) for _ in range(ndim))
        shape = tuple(draw(st.integers(1, 13)) for _ in range(ndim))
        sharded = draw(st.booleans())
        dtype = draw(st.sampled_from(["uint8", "uint16", "float32", "int64"]))
        writes = []
        for _ in range(draw(st.integers(1, 4))):
            region = []
            for n in shape:
                a = draw(st.integers(0, n - 1))
                region.append(slice(a, draw(st.integers(a + 1, n))))
            writes.append((tuple(region), draw(st.integers(0, 2 ** 31 - 1))))
        return shape, chunks, tuple(c * m for c, m in zip(chunks, mult)) if sharded else None, dtype, writes

    @hyp.settings(max_examples=40, deadline=None, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(stores())
    def run(case):
        shape, chunks, shards, dtype, writes = case
        path = tmp_path_factory.mktemp("prop") / "s.zarr"
        arr = zarr3.open(path, mode="w").create_array(name="0", shape=shape, chunks=chunks, shards=shards, dtype=dtype)
        model = np.zeros(shape, dtype)
        for region, seed in writes:
            block = (np.random.default_rng(seed).random([r.stop - r.start for r in region]) * 3).astype(dtype)
            arr[region] = block                       # mostly small values: some chunks stay equal to the fill value
            model[region] = block
        back = zarr3.open(path, mode="r")["0"]
        assert np.array_equal(back[...], model)
        for region, _ in writes:
            assert np.array_equal(back[region], model[region])
        staged = back.read_chunk_major()
        other = zarr3.open(path, mode="r+").create_array(name="1", shape=shape, chunks=chunks, shards=shards,
                                                         dtype=dtype)
        other.write_chunk_major(staged)
        assert np.array_equal(other[...], model)

    run()


def test_bulk_path_layout_rules():
    """Host-side argument rules of the device staging kernels (no GPU needed): which arrays take the bulk path and how
    a voxel is sized inside the volume and inside a chunk."""
    from interactive_unet_b200.engine import Engine
    assert Engine._voxel_layout((512, 512, 512, 2), (128, 128, 128, 2), 1) == (512, 512, 512, 2, 2, 128, 128, 128)
    assert Engine._voxel_layout((256, 256, 256, 1), (128, 128, 128, 2), 1) == (256, 256, 256, 1, 2, 128, 128, 128)
    assert Engine._voxel_layout((40, 36, 44), (16, 16, 16), 2) == (40, 36, 44, 2, 2, 16, 16, 16)
    assert Engine._voxel_layout((8, 8, 8, 3), (4, 4, 4, 3), 4) == (8, 8, 8, 12, 12, 4, 4, 4)
    for shape, chunks in [((8, 8, 8, 4), (4, 4, 4, 2)), ((8, 8), (4, 4)), ((8, 8, 8), (4, 4)), ((8, 8, 8, 2, 2), (4, 4, 4, 2, 4))]:
        with pytest.raises(ValueError):
            Engine._voxel_layout(shape, chunks, 1)

    class A:
        def __init__(self, shape, chunks):
            self.shape, self.chunks, self.ndim = shape, chunks, len(shape)
            self.chunk_grid = tuple(-(-n // c) for n, c in zip(shape, chunks))
    assert iu_utils._chunked_on_three_axes(A((512, 512, 512, 2), (128, 128, 128, 2)))
    assert iu_utils._chunked_on_three_axes(A((256, 256, 256, 1), (128, 128, 128, 2)))      # pyramid level, class axis halved
    assert iu_utils._chunked_on_three_axes(A((40, 36, 44), (16, 16, 16)))
    assert not iu_utils._chunked_on_three_axes(A((8, 8, 8, 4), (4, 4, 4, 2)))              # class axis chunked
    assert not iu_utils._chunked_on_three_axes(A((8, 8), (4, 4)))

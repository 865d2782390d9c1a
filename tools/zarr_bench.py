"""Stage timings of `predict_volumes` on disk: Zarr v3 store in -> prediction store + pyramid out.

    python tools/zarr_bench.py [--edge 512] [--input-size 512] [--classes 2] [--dir /tmp/iu_zarr_bench]

Prints one JSON line: wall time of every stage (store read, H2D + chunk scatter, prediction, chunk gather + D2H,
compress + write, pyramid), the two staging kernels against the HBM roofline (algorithmic bytes = one read + one write
of the level), and -- as the CPU baseline of the staging step -- the same chunk re-ordering done with numpy slicing on
the host, which is what a Zarr library does inside `array[...] = data` / `array[...]`.
Synthetic volume: blocky noise (8-voxel cells + 3 bits of per-voxel noise), so zstd sees realistic redundancy.
"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactive_unet_b200 as iu  # noqa: E402
from interactive_unet_b200 import utils, zarr3  # noqa: E402


def synthetic_volume(edge, seed=1):
    rng = np.random.default_rng(seed)
    cells = rng.integers(0, 224, (edge // 8,) * 3, dtype=np.uint8)
    vol = np.repeat(np.repeat(np.repeat(cells, 8, 0), 8, 1), 8, 2)
    return vol + rng.integers(0, 8, vol.shape, dtype=np.uint8)


def device_ms(fn, stream_sync, reps=5):
    fn()
    stream_sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    stream_sync()
    return 1e3 * (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=512)
    ap.add_argument("--input-size", type=int, default=512)
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--dir", default="/tmp/iu_zarr_bench")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n, c = args.edge, args.classes
    shutil.rmtree(args.dir, ignore_errors=True)
    os.makedirs(os.path.join(args.dir, "data", "image_volumes"))
    os.makedirs(os.path.join(args.dir, "data", "predicted_volumes"))
    os.chdir(args.dir)
    vol = synthetic_volume(n)
    t = {}
    t0 = time.perf_counter()
    utils.create_multiscale_zarr(vol, "data/image_volumes/vol.zarr")
    t["create_input_store_s"] = time.perf_counter() - t0

    torch.manual_seed(1234)
    model = iu.UNet(num_classes=c).to(dev).eval()
    eng = utils._engine(dev)
    sync = lambda: torch.cuda.synchronize()

    # -- the stages of predict_volumes, timed one by one (second pass: page cache warm, allocations done)
    for rep in range(2):
        arr = zarr3.open("data/image_volumes/vol.zarr", mode="r")["0"]
        t0 = time.perf_counter()
        buf, staged = utils._pinned.acquire(arr.chunk_major_shape(), torch.uint8)    # as read_array_to_device does
        arr.read_chunk_major(out=staged.numpy())
        t1 = time.perf_counter()
        vol_dev = eng.from_chunks(staged.to(dev, non_blocking=True), arr.shape, arr.chunks)
        sync()
        utils._pinned.release(buf)
        t2 = time.perf_counter()
        out_dev = iu.predict.predict_volume_array(model, vol_dev, input_size=args.input_size, num_classes=c)
        sync()
        t3 = time.perf_counter()
        root = zarr3.open("data/predicted_volumes/vol.zarr", mode="w")
        final = root.create_array(name="0", shape=list(arr.shape) + [c], chunks=(128, 128, 128, c),
                                  shards=(256, 256, 256, c), dtype="uint8", overwrite=True)
        staged_dev = eng.to_chunks(out_dev, final.chunks)
        buf, host = utils._pinned.acquire(staged_dev.shape, torch.uint8)             # as write_array_from_device does
        host.copy_(staged_dev)
        sync()
        t4 = time.perf_counter()
        final.write_chunk_major(host.numpy())
        utils._pinned.release(buf)
        t5 = time.perf_counter()
        utils.add_multiscales("data/predicted_volumes/vol.zarr", scale=0.5, level0=out_dev)
        t6 = time.perf_counter()
    t.update(read_decompress_s=t1 - t0, h2d_scatter_s=t2 - t1, predict_s=t3 - t2, gather_d2h_s=t4 - t3,
             compress_write_s=t5 - t4, pyramid_s=t6 - t5, total_s=t6 - t0)
    assert torch.equal(vol_dev.cpu(), torch.from_numpy(vol))
    assert np.array_equal(zarr3.open("data/predicted_volumes/vol.zarr")["0"][:64, :64], out_dev[:64, :64].cpu().numpy())

    # -- the whole drop-in call, as the GUI makes it
    t0 = time.perf_counter()
    iu.predict.predict_volumes(input_size=args.input_size, num_classes=c)
    t["predict_volumes_call_s"] = time.perf_counter() - t0
    # ... and with three volumes queued: the next store is decompressed while this one is on the device, and the
    # previous one's shards are still being compressed (steady-state cost per volume)
    for k in (2, 3):
        shutil.copytree("data/image_volumes/vol.zarr", f"data/image_volumes/vol{k}.zarr")
    t0 = time.perf_counter()
    iu.predict.predict_volumes(input_size=args.input_size, num_classes=c)
    t["predict_volumes_3_volumes_s"] = time.perf_counter() - t0
    for k in (2, 3):
        assert np.array_equal(zarr3.open(f"data/predicted_volumes/vol{k}.zarr")["1"][...],
                              zarr3.open("data/predicted_volumes/vol.zarr")["1"][...])
        shutil.rmtree(f"data/predicted_volumes/vol{k}.zarr")
    stored = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk("data/predicted_volumes") for f in fs)

    # -- staging kernels vs the HBM roofline
    peaks_file = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = 6535.1
    try:
        peak = float(json.load(open(peaks_file))["hbm_gbs"])
    except Exception:
        pass
    # (timed on a 1 GiB array so that the ~25 us of call + synchronise per launch stays below 10 % of the kernel)
    k = {}
    big = torch.randint(0, 256, (1024, 1024, 512, c), dtype=torch.uint8, device=dev)
    big_staged = eng.to_chunks(big, final.chunks)
    ms = device_ms(lambda: eng.to_chunks(big, final.chunks, out=big_staged), sync)
    k["to_chunks"] = dict(ms=ms, bytes=2 * big.numel(), gbps=2 * big.numel() / ms / 1e6)
    ms = device_ms(lambda: eng.from_chunks(big_staged, big.shape, final.chunks, out=big), sync)
    k["from_chunks"] = dict(ms=ms, bytes=2 * big.numel(), gbps=2 * big.numel() / ms / 1e6)
    tables = utils.zoom_tables(big.shape, tuple(int(x * 0.5) for x in big.shape), 0.5, 256)
    lvl1 = torch.empty([x.size for x in tables], dtype=torch.uint8, device=dev)
    ms = device_ms(lambda: eng.zoom_nearest(big, tables, out=lvl1), sync)
    # algorithmic bytes: every output byte written once, and the 32-byte sectors that hold its source bytes read
    # once: every other row and plane of the source, i.e. a quarter of it
    zb = lvl1.numel() + big.numel() // 4
    k["zoom_nearest"] = dict(ms=ms, bytes=zb, gbps=zb / ms / 1e6)
    del big, big_staged, lvl1
    for v in k.values():
        v["frac_of_hbm_peak"] = v["gbps"] / peak

    # -- CPU baseline of the staging step: the same re-ordering with numpy slicing, one thread (what zarr does)
    out_host = out_dev.cpu().numpy()
    t0 = time.perf_counter()
    staged_np = np.zeros(final.chunk_major_shape(), np.uint8)
    for i, (gz, gy, gx) in enumerate(np.ndindex(*final.chunk_grid[:3])):
        p = out_host[gz * 128:(gz + 1) * 128, gy * 128:(gy + 1) * 128, gx * 128:(gx + 1) * 128]
        staged_np[i, :p.shape[0], :p.shape[1], :p.shape[2]] = p
    cpu_s = time.perf_counter() - t0
    assert np.array_equal(staged_np, staged_dev.cpu().numpy())

    print(json.dumps({"workload": f"{n}^3 uint8 volume, {c} classes, input_size {args.input_size}, chunks 128 / shards 256",
                      "host_cores": os.cpu_count(), "stages": {a: round(b, 4) for a, b in t.items()},
                      "voxels_per_s_disk_to_disk": 3 * n ** 3 / t["predict_volumes_3_volumes_s"],
                      "stored_bytes_prediction_store": stored, "hbm_peak_gbps": peak,
                      "kernels": {a: {x: (round(y, 4) if isinstance(y, float) else y) for x, y in b.items()}
                                  for a, b in k.items()},
                      "cpu_baseline_chunk_reorder": {"seconds": round(cpu_s, 4), "cores": 1,
                                                     "gbps": 2 * out_host.size / cpu_s / 1e9}}))


if __name__ == "__main__":
    main()

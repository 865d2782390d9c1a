"""Drop-in for the prediction entry points of `/root/reference/interactive_unet/predict.py`.

`predict_slice` (`:16-47`), `find_max_batch_size` (`:49-77`), `predict_block` (`:79-112`) and
`predict_volumes` (`:114-266`) keep the reference's signatures, return types and on-disk side effects;
the arithmetic between "uint8 volume" and "uint8 probabilities / labels" runs in the native sm_100a
engine (`libiunet_b200.so`): slice gather + normalise, the U-Net, per-slice softmax, cross-axis
accumulate / average, Gaussian-window blend, uint8 quantise and argmax.  No stage falls back to
PyTorch or the CPU; if the library or the GPU is missing these functions raise.

`predict_volume_array` is the in-memory core that `predict_volumes` calls per Zarr volume.
"""
import glob
import os
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import unet
from .engine import gaussian_window_1d

# utils.COLORS of the reference (utils.py:304-306): class i is drawn with COLORS[i + 1]
COLORS = np.array([[0, 0, 0], [230, 25, 75], [60, 180, 75], [255, 225, 25], [0, 130, 200], [245, 130, 48],
                   [145, 30, 180], [70, 240, 240], [240, 50, 230], [210, 245, 60], [170, 255, 195]], dtype=np.uint8)

_model_cache = {}
_model_cache_lock = threading.Lock()


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("interactive_unet_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch.device('cuda')


def _load_model(num_channels, num_classes, device):
    """`predict.py:21-27,127-135`: the checkpoint if present, else a fresh model.  The reference reloads
    the file on every call; here the loaded model (and its engine-resident weights) is cached and
    refreshed when the trainer rewrites the file (`trainer.py:42-49`)."""
    model_path = os.path.join('model', 'model.ckpt')
    if os.path.isfile(model_path):
        st = os.stat(model_path)
        key = (os.path.abspath(model_path), st.st_mtime_ns, st.st_size, str(device))
        with _model_cache_lock:
            model = _model_cache.get(key)
            if model is None:
                try:
                    model = unet.UNet.load_from_checkpoint(checkpoint_path=model_path).to(device)
                except NotImplementedError as e:
                    raise NotImplementedError(
                        f"{model_path} was trained with a configuration the B200 engine does not accelerate ({e}). "
                        "Keep `from interactive_unet import predict` for this model, or train with "
                        "architecture='U-Net' and encoder_name='resnet34' / 'resnet18' (INTEGRATION.md section 1).") from e
                model.eval()
                _model_cache.clear()
                _model_cache[key] = model
        return model
    model = unet.UNet(num_channels=num_channels, num_classes=num_classes).to(device)
    model.eval()
    return model


def categorical_to_colored(mask):
    """`utils.py:351-357`."""
    colored = np.zeros((mask.shape[0], mask.shape[1], 3), dtype='uint8')
    for i in range(mask.shape[-1]):
        colored[mask[:, :, i] == 255, :] = COLORS[i + 1]
    return colored


def predict_slice(image_slice, num_channels=1, num_classes=2, return_probabilities=False):
    """`predict.py:16-47`: uint8 `[H,W]` -> colour overlay uint8 `[H,W,3]` (or probabilities `[1,H,W,C]`)."""
    device = _require_cuda()
    model = _load_model(num_channels, num_classes, device)
    x = (image_slice[None, None, :, :] / 255).astype('float32')
    with torch.inference_mode():
        y_prob = model(torch.from_numpy(np.ascontiguousarray(x)).to(device)).cpu().numpy()
    y_prob = np.moveaxis(y_prob, 1, -1)
    y_pred = np.argmax(y_prob[0, :, :, :num_classes], axis=-1)
    y_pred = np.stack([y_pred == i for i in range(num_classes)], -1)
    y_pred = (y_pred * 255).astype('uint8')
    y_pred = categorical_to_colored(y_pred)
    return y_prob if return_probabilities else y_pred


def find_max_batch_size(model, input_size=256, start=4, max_limit=512):
    """`predict.py:49-77`: the largest power-of-two multiple of `start` (<= max_limit) whose activations
    fit.  The engine allocates its workspace up front, so this is a capacity query plus one real
    allocation probe per size instead of timed trial forwards."""
    batch_size, best = start, start
    device = model.device
    eng = model.engine()
    while batch_size <= max_limit:
        try:
            with torch.inference_mode(), eng.limit_batch(batch_size):
                test_batch = torch.zeros((batch_size, 1, input_size, input_size), dtype=torch.float32, device=device)
                _ = model(test_batch)
            best = batch_size
            batch_size *= 2
            torch.cuda.empty_cache()
        except RuntimeError as e:
            if "out of memory" in str(e):                                # predict.py:67-72
                torch.cuda.empty_cache()
                break
            raise
    eng.release_workspace()              # the probe's plans (one per size tried) go back to the driver
    torch.cuda.empty_cache()
    return best


def predict_block(model, block, num_classes=2, batch_size=8, axes=[0, 1, 2]):
    """`predict.py:79-112`: fp32 block `[S,S,S]` (values = uint8/255) -> mean probabilities fp32 `[S,S,S,C]`."""
    block = block.detach() if isinstance(block, torch.Tensor) else torch.as_tensor(np.asarray(block))
    size = block.shape[0]
    if tuple(block.shape) != (size, size, size):
        raise ValueError("predict_block expects a cubic block (predict.py:81)")
    eng = model.engine()
    if eng.num_classes != num_classes:
        raise ValueError(f"model has {eng.num_classes} classes, num_classes={num_classes} requested")
    out = np.empty((size, size, size, num_classes), dtype=np.float32)
    vol = block.to(torch.float32).contiguous()
    vol = vol if vol.is_cuda else vol.numpy()
    with eng.limit_batch(batch_size):
        eng.predict_volume(vol, axes=list(axes), window=None, out_mean=out)
    return out


def gaussian_3d(input_size, sigma=0.125, eps=1e-3):
    """`predict.py:327-347` (host helper, same values): the 3-D blending window of one block."""
    sigma = sigma * input_size
    coords = np.arange(input_size, dtype=np.float32) - (input_size - 1) / 2.0
    g = np.exp(-(coords ** 2) / (2 * sigma ** 2)).astype(np.float32)
    g /= g.max()
    gaussian = g[:, None, None] * g[None, :, None] * g[None, None, :]
    gaussian /= gaussian.max()
    return np.clip(gaussian, max(gaussian.min(), eps), 1.0)


def get_block_coordinates(volume_shape, input_size=256, overlap=0.25):
    """`predict.py:362-411`: (block_coords, padded_block_coords, local_block_coords), one row of six ints per block,
    blocks in the reference's nested (i, j, k) order with the padding centred on the volume."""
    volume_shape = np.asarray(volume_shape)
    step = input_size - overlap * input_size
    blocks_per_axis = np.ceil((volume_shape - overlap * input_size) / step).astype(int)
    padded_volume_shape = np.round(blocks_per_axis * input_size - (blocks_per_axis - 1) * input_size * overlap).astype(int)
    shift = (padded_volume_shape - volume_shape) // 2
    shift6 = np.concatenate([shift, shift])
    block_coords, padded_block_coords, local_block_coords = [], [], []
    for i in range(blocks_per_axis[0]):
        for j in range(blocks_per_axis[1]):
            for k in range(blocks_per_axis[2]):
                lo = np.array([i, j, k]) * input_size * (1 - overlap)
                padded = (np.concatenate([lo, lo + input_size]) - shift6).astype(int)
                clipped = np.concatenate([np.maximum(padded[:3], 0), np.minimum(padded[3:], volume_shape)])
                padded_block_coords.append(padded)
                block_coords.append(clipped)
                local_block_coords.append(clipped - np.concatenate([padded[:3], padded[:3]]))
    return np.array(block_coords), np.array(padded_block_coords), np.array(local_block_coords)


def get_padded_block(volume, i0, j0, k0, i1, j1, k1):
    """`predict.py:291-316`: the box clipped to the volume and padded back with numpy's 'reflect' mode."""
    shape = volume.shape
    before = [max(0, -i0), max(0, -j0), max(0, -k0)]
    after = [max(0, i1 - shape[0]), max(0, j1 - shape[1]), max(0, k1 - shape[2])]
    block = volume[max(i0, 0):min(i1, shape[0]), max(j0, 0):min(j1, shape[1]), max(k0, 0):min(k1, shape[2])]
    return np.pad(block, tuple(zip(before, after)), mode='reflect')


def get_shard_coordinates(volume_shape, shard_size=128):
    """`predict.py:318-325`."""
    starts = [np.arange(0, s, shard_size) for s in volume_shape]
    c = np.stack(np.meshgrid(*starts, indexing='ij'), -1).reshape(-1, 3)
    return np.concatenate([c, np.minimum(c + shard_size, volume_shape)], axis=1)


def tiled_device_bytes(shape, input_size, num_classes, n_axes=3, volume_on_device=True, out_on_device=True):
    """Device bytes the tiled mode holds for a `[D,H,W]` volume beyond what the caller already keeps there: the fp32
    `pred` / `weight` accumulators of `predict.py:181-198` (on disk in the reference; here a ring of `input_size` z
    planes in HBM -- finished z ranges are normalised and handed out while later blocks are still being predicted), the
    uint8 outputs and the volume itself when they arrive from / go to the host, one block and its per-axis probabilities."""
    vox = int(np.prod(shape, dtype=np.int64))
    bvox = int(input_size) ** 3
    ring = min(int(shape[0]), int(input_size)) * int(shape[1]) * int(shape[2])
    total = ring * (4 * num_classes + 4) + bvox * (1 + 4 * num_classes * n_axes)
    if not volume_on_device:
        total += vox
    if not out_on_device:
        total += vox * (num_classes + 1)
    return total


def _check_tiled_fits(eng, shape, input_size, num_classes, n_axes, volume_on_device, out_on_device):
    """The reference streams blocks through on-disk accumulators and handles any volume size; this engine keeps a ring
    of `input_size` planes of them in HBM plus the uint8 volume and result.  Fail before allocating anything, with the
    numbers, rather than with an allocator error half way."""
    need = tiled_device_bytes(shape, input_size, num_classes, n_axes, volume_on_device, out_on_device)
    need += eng.workspace_bytes(eng.auto_batch(input_size, input_size, input_size), input_size, input_size)
    free, _total = torch.cuda.mem_get_info(eng.device)
    avail = free + eng.held_bytes()                      # pooled scratch and cached plans are reused or released
    if need > avail:
        raise RuntimeError(
            f"CUDA out of memory: predicting a {shape[0]}x{shape[1]}x{shape[2]} volume in the tiled mode keeps "
            f"{need / 2**30:.1f} GiB on the device (the uint8 volume and result, and {4 * num_classes + 4} B per voxel "
            f"of fp32 accumulators for {min(shape[0], input_size)} z planes), {avail / 2**30:.1f} GiB are available; "
            f"split the volume (for example into z ranges that overlap by input_size * overlap) and predict the parts "
            f"separately")


def predict_volume_array(model, volume, input_size=None, num_classes=2, overlap=0.25, batch_size=None,
                         axes=[0, 1, 2], return_labels=False, out=None, out_labels=None):
    """In-memory core of `predict_volumes` (`predict.py:153,201,235-256`): uint8 volume `[D,H,W]` (numpy, or a CUDA
    tensor to keep everything device-resident) -> uint8 probabilities `[D,H,W,C]` (and uint8 argmax labels `[D,H,W]`
    with `return_labels=True`), same container kind as the input.  A cubic volume of edge `input_size` is one block
    (the fused single-block path); anything else is tiled into `input_size`^3 blocks with `overlap`, reflect padding
    and Gaussian blending exactly as the reference does, all on the device.
    `out` / `out_labels` may be preallocated (e.g. pinned host arrays) to avoid per-call allocation."""
    shape = tuple(int(v) for v in volume.shape)
    if len(shape) != 3:
        raise ValueError("predict_volume_array expects a 3-D uint8 volume")
    input_size = shape[0] if input_size is None else int(input_size)
    if input_size % 32:
        raise RuntimeError(f"Wrong input shape height={input_size}, width={input_size}. Expected image height and width "
                           f"divisible by 32.")
    eng = model.engine()
    if eng.num_classes != num_classes:
        raise ValueError(f"model has {eng.num_classes} classes, num_classes={num_classes} requested")
    window = gaussian_window_1d(input_size, sigma=0.125)               # predict.py:153
    lab = out_labels
    if isinstance(volume, torch.Tensor):
        volume = volume.contiguous()
        if out is None:
            out = torch.empty(shape + (num_classes,), dtype=torch.uint8, device=volume.device)
        if lab is None and return_labels:
            lab = torch.empty(shape, dtype=torch.uint8, device=volume.device)
    else:
        volume = np.ascontiguousarray(volume)
        if out is None:
            out = np.empty(shape + (num_classes,), dtype=np.uint8)
        if lab is None and return_labels:
            lab = np.empty(shape, dtype=np.uint8)
    with eng.limit_batch(batch_size):
        if shape == (input_size,) * 3:
            eng.predict_volume(volume, axes=list(axes), window=window, out_u8=out, out_labels=lab)
        else:
            if volume.dtype not in (np.uint8, torch.uint8):
                raise TypeError("the tiled mode reads uint8 volumes (predict.py:237)")
            _check_tiled_fits(eng, shape, input_size, num_classes, len(axes), isinstance(volume, torch.Tensor),
                              isinstance(out, torch.Tensor))
            _, padded, _ = get_block_coordinates(np.array(shape), input_size=input_size, overlap=overlap)
            eng.predict_tiled(volume, input_size, padded[:, :3], axes=list(axes), window=window, out_u8=out, out_labels=lab)
    return (out, lab) if return_labels else out


def predict_volumes(input_size=256, num_channels=1, num_classes=2, overlap=0.25, chunk_size=128, shard_size=256,
                    batch_size=None, axes=[0, 1, 2]):
    """`predict.py:114-266`: predict every `data/image_volumes/*.zarr` (level '0', uint8) into
    `data/predicted_volumes/<same name>` -- level '0' uint8 `[D,H,W,C]` with chunks `(chunk_size,)*3 + (C,)` and shards
    `(shard_size,)*3 + (C,)`, then the multiscale pyramid (`utils.add_multiscales`, `predict.py:261`).

    Data flow: host threads decompress the store's inner chunks into pinned memory -> one H2D copy -> the device puts
    them in `[D,H,W]` order -> prediction (single block, or tiled + blended) entirely on the device -> the device
    re-orders the uint8 result into inner chunks and zooms each pyramid level -> one D2H copy per level -> host threads
    compress and write one shard file each.  The reference's float32 `temp/pred.zarr` / `temp/weight.zarr` accumulators
    (`predict.py:181-198`) live in HBM instead, so no `temp/` directory is created (or removed, `predict.py:259`)."""
    import signal
    from . import utils, zarr3
    from .distributed import volumes_for_rank

    if threading.current_thread() is threading.main_thread():
        def handle_sigint(sig, frame):                       # predict.py:118-122
            print("\nCaught Ctrl+C \u2192 force exit")
            os._exit(1)
        signal.signal(signal.SIGINT, handle_sigint)

    device = _require_cuda()
    model = _load_model(num_channels, num_classes, device)
    # one process per GPU under torchrun: every rank takes its share of the (independent) volumes, no collective
    volume_files = volumes_for_rank(np.sort(glob.glob('data/image_volumes/*.zarr')))

    def open_and_stage(f):
        volume = zarr3.open(f, mode='r')['0']                # highest resolution (predict.py:167)
        if volume.dtype != np.uint8 or volume.ndim != 3:
            raise TypeError(f"{f}: level '0' must be a 3-D uint8 array (predict.py:237 divides by 255), "
                            f"got {volume.dtype} {volume.shape}")
        return utils.stage_array(volume)

    def finish(job):
        """Wait for a volume's compression / file tasks and report it."""
        if job is not None:
            name, shape, start_time, pending = job
            for p in pending:
                p.result()
            print(f'Completed volume {name} {shape} in {time.time() - start_time}.')

    # Three volumes are in flight: i + 1 is being decompressed into pinned memory by a helper thread, i is on the
    # device, and the shards of i - 1 are still being compressed and written by the pool.
    prefetch = ThreadPoolExecutor(max_workers=1, thread_name_prefix="iu-prefetch")
    staged_next = prefetch.submit(open_and_stage, volume_files[0]) if len(volume_files) else None
    previous = None
    try:
        for n, f in enumerate(volume_files):
            start_time = time.time()
            staged = staged_next.result()
            staged_next = prefetch.submit(open_and_stage, volume_files[n + 1]) if n + 1 < len(volume_files) else None
            shape = tuple(staged.arr.shape)
            save_path = f.replace('image_volumes', 'predicted_volumes')
            root = zarr3.open(save_path, mode='w')
            final_predictions = root.create_array(name='0', shape=list(shape) + [num_classes],
                                                  chunks=(chunk_size, chunk_size, chunk_size, num_classes),
                                                  shards=(shard_size, shard_size, shard_size, num_classes),
                                                  dtype='uint8', overwrite=True)
            print(f'\nSegmenting {os.path.basename(f)}...')
            volume_dev = staged.to_device(device)
            out_dev = predict_volume_array(model, volume_dev, input_size=input_size, num_classes=num_classes,
                                           overlap=overlap, batch_size=batch_size, axes=axes)
            del volume_dev
            finish(previous)
            previous = None
            print('Postprocessing and generating multiscale pyramid...')
            pending = [utils.write_array_from_device(final_predictions, out_dev, wait=False)]
            previous = (os.path.basename(f), shape, start_time, pending)
            utils.add_multiscales(save_path, scale=0.5, level0=out_dev, _defer=pending)
            del out_dev
    finally:
        try:
            finish(previous)
        finally:
            if staged_next is not None:
                try:
                    staged_next.result().release()
                except Exception:
                    pass
            prefetch.shutdown(wait=True)
            # the engine is cached with the model: give its workspace back (the trainer shares this GPU)
            model.engine().release_workspace()
    print('\nAll volumes segmented.\n')

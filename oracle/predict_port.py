"""numpy restatement of the arithmetic in the reference's `predict.py`.  TEST INFRASTRUCTURE.

Every function cites the lines of `/root/reference/interactive_unet/predict.py`
it follows.  Pinned against the verbatim reference (see
`oracle/reference_loader.py`) by `tests/test_oracle.py` in the build container
and against `tests/golden/*.npz` everywhere else.

Conventions: volumes are C-order `[z, y, x]`; a slice along
axis `a` is `moveaxis(block, a, 0)[i]`, i.e. image (y,x) for a=0, (z,x) for
a=1, (z,y) for a=2; all tail arithmetic is IEEE fp32 in the reference's order.
"""
import numpy as np


def normalise_u8(u8):
    """`predict.py:237`  `block.astype('float32') / 255.0`  (true fp32 division)."""
    return u8.astype(np.float32) / np.float32(255.0)


def slice_batch(block, axis, start, count):
    """`predict.py:91,95`  slices `start..start+count` of `block` along `axis`, as [B,1,S,S]."""
    moved = np.moveaxis(block, axis, 0)
    return np.ascontiguousarray(moved[start:start + count])[:, None]


def scatter_batch(acc, probs_bhwc, axis, start):
    """`predict.py:101-106`  add a batch of per-slice probabilities `[B,S,S,C]` into `[z,y,x,C]`."""
    b = probs_bhwc.shape[0]
    if axis == 0:
        acc[start:start + b] += probs_bhwc
    elif axis == 1:
        acc[:, start:start + b] += probs_bhwc.transpose(1, 0, 2, 3)
    elif axis == 2:
        acc[:, :, start:start + b] += probs_bhwc.transpose(1, 2, 0, 3)
    else:
        raise ValueError("axis must be 0, 1 or 2")


def predict_block(forward, block, num_classes=2, batch_size=8, axes=(0, 1, 2)):
    """`predict.py:79-112`.  `forward(f32[B,1,S,S]) -> f32[B,C,S,S]` probabilities."""
    block = np.asarray(block, dtype=np.float32)
    size = block.shape[0]                                     # :81 cubic blocks only
    acc = np.zeros((size, size, size, num_classes), np.float32)   # :85
    for axis in axes:                                         # :87
        for start in range(0, size, batch_size):              # :93
            x = slice_batch(block, axis, start, batch_size)
            p = np.asarray(forward(x), dtype=np.float32)
            scatter_batch(acc, np.moveaxis(p, 1, -1), axis, start)   # :98
    acc /= np.float32(len(axes))                              # :110 true division
    return acc


def gaussian_1d(size, sigma=0.125):
    """The 1-D factor of `gaussian_3d`, `predict.py:333-339`."""
    s = sigma * size
    coords = np.arange(size, dtype=np.float32) - (size - 1) / 2.0
    g = np.exp(-(coords ** 2) / (2 * s ** 2)).astype(np.float32)
    g /= g.max()
    return g


def gaussian_3d(size, sigma=0.125, eps=1e-3):
    """`predict.py:327-347`."""
    g = gaussian_1d(size, sigma)
    w = g[:, None, None] * g[None, :, None] * g[None, None, :]
    w /= w.max()
    return np.clip(w, max(w.min(), eps), 1.0)


def block_coordinates(volume_shape, input_size=256, overlap=0.25):
    """`predict.py:362-411`: (clipped, padded, local) corner lists, block order i, j, k."""
    shape = np.asarray(volume_shape)
    step = input_size * (1 - overlap)
    nblk = np.ceil((shape - overlap * input_size) / (input_size - overlap * input_size)).astype(int)
    padded_shape = np.round(nblk * input_size - (nblk - 1) * input_size * overlap).astype(int)
    shift = (padded_shape - shape) // 2
    clipped, padded, local = [], [], []
    for i in range(nblk[0]):
        for j in range(nblk[1]):
            for k in range(nblk[2]):
                lo = np.array([i, j, k]) * step
                c = (np.concatenate([lo, lo + input_size]) - np.concatenate([shift, shift])).astype(int)
                padded.append(c)
                lo_c = np.maximum(c[:3], 0)
                hi_c = np.minimum(c[3:], shape)
                clipped.append(list(lo_c) + list(hi_c))
                local.append(list(lo_c - c[:3]) + list(hi_c - c[:3]))
    return np.array(clipped), np.array(padded), np.array(local)


def padded_block(volume, i0, j0, k0, i1, j1, k1):
    """`predict.py:291-316` (the second definition, which shadows `:281-289`)."""
    shp = volume.shape
    lo, hi = (i0, j0, k0), (i1, j1, k1)
    sl = tuple(slice(max(a, 0), min(b, n)) for a, b, n in zip(lo, hi, shp))
    pad = tuple((max(0, -a), max(0, b - n)) for a, b, n in zip(lo, hi, shp))
    return np.pad(volume[sl], pad_width=pad, mode="reflect")


def shard_coordinates(volume_shape, shard_size=128):
    """`predict.py:318-325`."""
    starts = [np.arange(0, s, shard_size) for s in volume_shape]
    lo = np.stack(np.meshgrid(*starts, indexing="ij"), -1).reshape(-1, 3)
    return np.concatenate([lo, np.minimum(lo + shard_size, volume_shape)], axis=1)


def quantise(pred, weight, eps=1e-3):
    """`predict.py:255`  `(255 * pred / max(weight, eps)[..., None]).astype('uint8')`."""
    return (255 * pred / np.maximum(weight, eps)[..., None]).astype("uint8")


def blend_single_block(mean_probs, window):
    """`predict.py:244-245` for the one-block case (N == input_size): returns (pred, weight)."""
    return mean_probs * window[..., None], window.copy()


def predict_volume(forward, volume_u8, input_size, num_classes=2, overlap=0.25, batch_size=8,
                   axes=(0, 1, 2)):
    """`predict.py:153,201,235-256` on an in-memory uint8 volume: uint8 `[Z,Y,X,C]`."""
    shape = np.asarray(volume_u8.shape)
    window = gaussian_3d(input_size)
    pred = np.zeros(tuple(shape) + (num_classes,), np.float32)
    weight = np.zeros(tuple(shape), np.float32)
    clipped, padded, local = block_coordinates(shape, input_size, overlap)
    for c, p, l in zip(clipped, padded, local):
        blk = normalise_u8(padded_block(volume_u8, *p))
        out = predict_block(forward, blk, num_classes, batch_size, axes)
        i0, j0, k0, i1, j1, k1 = c
        a0, b0, c0, a1, b1, c1 = l
        pred[i0:i1, j0:j1, k0:k1] += out[a0:a1, b0:b1, c0:c1, :] * window[a0:a1, b0:b1, c0:c1, None]
        weight[i0:i1, j0:j1, k0:k1] += window[a0:a1, b0:b1, c0:c1]
    return quantise(pred, weight)


def labels_from_probs(probs, num_classes):
    """`predict.py:38`  `np.argmax(y_prob[..., :num_classes], axis=-1)` (first maximum wins)."""
    return np.argmax(probs[..., :num_classes], axis=-1)


def slice_overlay(probs_1hwc, num_classes, palette):
    """`predict.py:37-42` + `utils.py:351-357`: argmax -> one-hot*255 -> colour image."""
    lab = labels_from_probs(probs_1hwc[0], num_classes)
    out = np.zeros(lab.shape + (3,), np.uint8)
    for i in range(num_classes):
        out[lab == i] = palette[i + 1]
    return out


def resize_volume(src, dst, scale=0.5, block_size=512, order=0):
    """`utils.py:29-48`: block-wise `scipy.ndimage.zoom` of `src` assigned into `dst` (numpy arrays; `dst` modified in
    place).  scipy is a pinned dependency of the reference (`pyproject.toml:20`) and supplies the zoom here too; the
    product derives the equivalent index tables itself (`utils.zoom_tables`)."""
    from scipy import ndimage
    n = src.shape
    for i in range(0, n[0], block_size):
        i0, i1 = i, min(i + block_size, n[0])
        for j in range(0, n[1], block_size):
            j0, j1 = j, min(j + block_size, n[1])
            for k in range(0, n[2], block_size):
                k0, k1 = k, min(k + block_size, n[2])
                dst[int(i0 * scale):int(i1 * scale), int(j0 * scale):int(j1 * scale), int(k0 * scale):int(k1 * scale)] = \
                    ndimage.zoom(src[i0:i1, j0:j1, k0:k1], scale, order=order)


def multiscale_levels(level0, chunk_shape, shard_shape, scale=0.5):
    """`utils.py:50-80` on in-memory arrays: the list of levels 1, 2, ... that `add_multiscales` creates from `level0`
    (raises where the reference raises, except its `UnboundLocalError` for zero steps: an empty list here)."""
    steps = int(np.floor(np.log((np.array(level0.shape) / np.array(chunk_shape)).max()) / np.log(1 / scale)))
    levels, cur = [], level0
    for _ in range(steps):
        nxt = np.zeros(tuple(int(x * scale) for x in cur.shape), cur.dtype)     # a fresh zarr array reads as fill 0
        resize_volume(cur, nxt, scale=scale, block_size=shard_shape[0], order=0)
        levels.append(nxt)
        cur = nxt
    return levels

"""CPU tests of the oracle itself: the numpy port against the golden vectors recorded from the
verbatim reference (`oracle/make_golden.py`), and against the verbatim reference when it is present."""
import os

import numpy as np
import pytest
import torch

from oracle import predict_port as pp
from oracle import reference_loader
from oracle.make_golden import toy_model_numpy

BLOCK_CASES = ["block_s16_c2_a012", "block_s16_c4_a012", "block_s16_c3_a20", "block_s8_c2_a1"]
VOLUME_CASES = ["volume_single_s32_c2", "volume_single_s32_c4", "volume_tiled_s32_c2", "volume_tiled_s16_c3"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", BLOCK_CASES)
def test_predict_block_matches_golden(golden_dir, name):
    g = _load(golden_dir, name)
    c = int(g["num_classes"])
    out = pp.predict_block(lambda x: toy_model_numpy(x, c), pp.normalise_u8(g["volume"]), c, int(g["batch_size"]),
                           tuple(int(a) for a in g["axes"]))
    assert out.dtype == np.float32
    assert np.array_equal(out, g["mean_probs"])


def test_predict_block_is_batch_size_independent(golden_dir):
    g = _load(golden_dir, "block_s16_c2_a012")
    f = lambda x: toy_model_numpy(x, 2)  # noqa: E731
    a = pp.predict_block(f, pp.normalise_u8(g["volume"]), 2, 3, (0, 1, 2))      # ragged last batch
    assert np.array_equal(a, g["mean_probs"])


def test_gaussian_window_matches_golden(golden_dir):
    g = _load(golden_dir, "gaussian3d.npz"[:-4])
    for s in (8, 16, 32, 48):
        assert np.array_equal(pp.gaussian_3d(s), g[f"w{s}"])
    w = pp.gaussian_3d(128)
    assert np.array_equal(np.stack([w[i, i, i] for i in range(128)]), g["diag128"])
    assert np.array_equal(w[63, 3, :], g["row128"])


def test_window_from_1d_factor_is_bit_exact():
    """The factorisation the CUDA reduce kernel uses: clip((g[z]*g[y])*g[x]/gmax, lo, 1)."""
    from interactive_unet_b200.engine import gaussian_window_1d
    for s in (8, 16, 32, 48, 96):
        g, gmax, lo = gaussian_window_1d(s)
        w = (g[:, None, None] * g[None, :, None]) * g[None, None, :]
        w = np.clip(w / np.float32(gmax), np.float32(lo), np.float32(1.0))
        assert np.array_equal(w, pp.gaussian_3d(s))


def test_block_coordinates_match_golden(golden_dir):
    g = _load(golden_dir, "coordinates")
    n = 0
    while f"case{n}_args" in g:
        a = g[f"case{n}_args"]
        c, p, l = pp.block_coordinates(tuple(a[:3]), int(a[3]), a[4] / 100.0)
        assert np.array_equal(c, g[f"case{n}_clipped"])
        assert np.array_equal(p, g[f"case{n}_padded"])
        assert np.array_equal(l, g[f"case{n}_local"])
        n += 1
    assert n >= 7
    assert np.array_equal(pp.shard_coordinates((100, 80, 60), 32), g["shards_100_80_60_32"])


def test_padded_block_matches_golden(golden_dir):
    g = _load(golden_dir, "padded_block")
    for i, box in enumerate(g["boxes"]):
        assert np.array_equal(pp.padded_block(g["volume"], *box), g[f"out{i}"])


@pytest.mark.parametrize("name", VOLUME_CASES)
def test_predict_volume_matches_golden(golden_dir, name):
    g = _load(golden_dir, name)
    c = int(g["num_classes"])
    out = pp.predict_volume(lambda x: toy_model_numpy(x, c), g["volume"], int(g["input_size"]), c,
                            axes=tuple(int(a) for a in g["axes"]))
    assert out.dtype == np.uint8 and out.shape == g["out_u8"].shape
    assert np.array_equal(out, g["out_u8"])


def test_normalise_is_true_division():
    u = np.arange(256, dtype=np.uint8)
    assert np.array_equal(pp.normalise_u8(u), (u / 255).astype("float32"))         # predict.py:30 form
    assert not np.array_equal(pp.normalise_u8(u), u.astype(np.float32) * np.float32(1 / 255))


def test_labels_first_maximum_wins():
    p = np.array([[0.5, 0.5, 0.0], [0.2, 0.4, 0.4], [0.1, 0.2, 0.7]], np.float32)
    assert pp.labels_from_probs(p, 3).tolist() == [0, 1, 2]


def test_network_restatement_shapes_and_work():
    from oracle.smp_unet_resnet34 import RefUNet, conv_macs_per_slice
    assert conv_macs_per_slice(64, 2) * 64 == 30882660352          # 30.883 GMAC at S=512
    m = RefUNet(1, 4).eval()
    with torch.inference_mode():
        y = m(torch.rand(2, 1, 64, 32))
    assert y.shape == (2, 4, 64, 32)
    assert torch.allclose(y.sum(1), torch.ones(2, 64, 32), atol=1e-5)
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 1, 48, 64))


def test_encoder_is_torchvision_resnet34():
    import torchvision
    from oracle.smp_unet_resnet34 import RefSmpUnetResnet34
    ours = RefSmpUnetResnet34(3, 2).eval()
    tv = torchvision.models.resnet34(weights=None).eval()
    tv.load_state_dict({k: v for k, v in ours.encoder.state_dict().items()}, strict=False)
    x = torch.rand(1, 3, 64, 64)
    with torch.inference_mode():
        f = ours.encode(x)[-1]
        t = tv.layer4(tv.layer3(tv.layer2(tv.layer1(tv.maxpool(tv.relu(tv.bn1(tv.conv1(x))))))))
    assert torch.equal(f, t)


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree only exists in the build container")
def test_port_matches_verbatim_reference():
    ref = reference_loader.load()
    from oracle.make_golden import ExactToyModel
    rng = np.random.default_rng(3)
    vol = rng.integers(0, 256, (16, 16, 16), dtype=np.uint8)
    want = ref.predict_block(ExactToyModel(3), torch.tensor(vol.astype("float32") / 255.0), num_classes=3,
                             batch_size=5, axes=[1, 2, 0])
    got = pp.predict_block(lambda x: toy_model_numpy(x, 3), pp.normalise_u8(vol), 3, 5, (1, 2, 0))
    assert np.array_equal(want, got)
    assert np.array_equal(ref.gaussian_3d(24), pp.gaussian_3d(24))
    for a, b in zip(ref.get_block_coordinates(np.array((50, 70, 90)), 32, 0.25), pp.block_coordinates((50, 70, 90), 32)):
        assert np.array_equal(a, b)


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("shape,chunk,shard", [((48, 32, 64, 2), 8, 16), ((40, 24, 56), 8, 16), ((34, 32, 32, 4), 8, 16)])
def test_pyramid_port_matches_verbatim_reference(shape, chunk, shard):
    """`predict_port.multiscale_levels` vs the verbatim `utils.add_multiscales` (`utils.py:50-80`) on fresh data,
    including a shape the reference fails on part-way (34 -> 17 -> 8: the error must be the same one)."""
    from oracle.make_golden import run_reference_add_multiscales
    rng = np.random.default_rng(11)
    vol = rng.integers(1, 255, shape, dtype=np.uint8)
    tail = shape[3:]
    want, err = run_reference_add_multiscales(vol, (chunk,) * 3 + tail, (shard,) * 3 + tail)
    try:
        got = pp.multiscale_levels(vol, (chunk,) * 3 + tail, (shard,) * 3 + tail)
    except ValueError as e:
        assert err is not None and err[0] == "ValueError" and err[1] == str(e)
    else:
        assert err is None and len(got) == len(want)
        for k, lv in enumerate(got):
            assert np.array_equal(lv, want[str(k + 1)])

"""CPU-only tests: the C-ABI library loads and exports what `include/iunet_b200.h` declares, the host
side fails loudly without a GPU, checkpoints in the reference's format load, and the multi-GPU host
logic (z-slab partition + all-to-all layout) is exercised with gloo, world_size 2."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "iunet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(iu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(built_library):
    import ctypes
    lib = ctypes.CDLL(built_library)
    names = _header_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/iunet_b200.h but not exported"
    lib.iu_abi_version.restype = ctypes.c_int
    assert lib.iu_abi_version() == 2


def test_header_is_plain_c_and_cxx(tmp_path):
    """The boundary is a C ABI: the header must compile on its own as C11 and as C++17 (no torch / CUDA types)."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "iunet_b200.h")
    for compiler, std, ext in (("gcc", "-std=c11", "c"), ("g++", "-std=c++17", "cpp")):
        if shutil.which(compiler) is None:
            pytest.skip(f"{compiler} not installed")
        src = tmp_path / f"use_header.{ext}"
        src.write_text('#include "iunet_b200.h"\nint main(void) { return sizeof(&iu_engine_predict_volume) == 0; }\n')
        r = subprocess.run([compiler, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                            "-I", os.path.dirname(hdr), str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_python_binding_mirrors_header(built_library):
    from interactive_unet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    _lib.load()


def test_product_package_never_touches_the_oracle():
    """`oracle/` is test infrastructure: no file of the shipped package (Python or CUDA) may import or name it,
    and none may read the reference tree."""
    pkg = os.path.join(ROOT, "interactive-unet_b200")
    for base, _, files in os.walk(pkg):
        if os.path.basename(base) in ("build", "__pycache__"):
            continue
        for name in files:
            if not name.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(base, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{name} imports oracle/"
            assert "/root/reference" not in text.replace("`/root/reference", "").replace("(`/root/reference", ""), \
                f"{name} reads the reference tree"


def test_bench_uses_the_oracle_only_in_its_cpu_arm():
    """`bench.py` may execute `oracle/` only in the CPU-baseline / reference-arm leg: every `oracle` import must sit
    inside `cpu_reference_sample` or `cpu_config0` (the two CPU timers both of those legs call), never at module level
    or in the GPU arm."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    where = []
    for fn in [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.Module))]:
        for node in (fn.body if isinstance(fn, ast.Module) else ast.walk(fn)):
            names = []
            if isinstance(node, ast.ImportFrom) and node.module:
                names = [node.module]
            elif isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            if any(n == "oracle" or n.startswith("oracle.") for n in names):
                where.append(getattr(fn, "name", "<module>"))
    assert where and set(where) == {"cpu_reference_sample", "cpu_config0"}, where


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(built_library):
    import interactive_unet_b200 as iu
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        iu.Engine(0)
    model = iu.UNet(num_classes=2).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.rand(1, 1, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        iu.predict.predict_slice(np.zeros((64, 64), np.uint8))


def test_unet_constructor_contract():
    import interactive_unet_b200 as iu
    m = iu.UNet(lr=1e-3, num_channels=1, num_classes=3, architecture='U-Net', encoder_name='resnet34')
    assert m.lr == 1e-3 and m.num_classes == 3
    assert all(k.startswith("model.") for k in m.state_dict())
    with pytest.raises(NotImplementedError):
        iu.UNet(architecture='U-Net++', encoder_name='resnet34')
    with pytest.raises(NotImplementedError):
        iu.UNet(encoder_name='mit_b0')
    assert isinstance(m.configure_optimizers(), torch.optim.AdamW)


def test_state_dict_matches_oracle_and_training_forward():
    import interactive_unet_b200 as iu
    from oracle.smp_unet_resnet34 import RefUNet
    ref = RefUNet(1, 2)
    m = iu.UNet(num_classes=2)
    assert list(m.state_dict()) == list(ref.state_dict())
    m.load_state_dict(ref.state_dict())
    m.train(), ref.train()
    x = torch.rand(2, 1, 64, 64)
    assert torch.equal(m(x), ref(x))                 # autograd path kept for the reference's trainer


def test_state_dict_layout_matches_smp_key_fixture(golden_dir):
    """Keys, order and shapes of the drop-in's `state_dict()` against the list assembled from torchvision's resnet34
    (encoder) and smp 0.5.0's published decoder / head (tests/golden/make_smp_keys.py): a reference checkpoint
    (`trainer.py:46-49`, keys `model.<smp key>`) loads with strict=True exactly when these agree."""
    import json
    import interactive_unet_b200 as iu
    want = json.load(open(os.path.join(golden_dir, "smp_unet_resnet34_keys.json")))["keys"]
    got = [[k, list(v.shape)] for k, v in iu.UNet(num_classes=2).state_dict().items()]
    assert got == [["model." + k, shp] for k, shp in want]


def test_trainer_hooks_fit_one_step():
    """`training_step` / `validation_step` / `_log_metrics` (unet.py:75-116) exist and one optimiser step runs through
    them with an injected metrics module (the reference's `interactive_unet.metrics` is used when it is importable)."""
    import interactive_unet_b200 as iu

    class Metrics:
        calls = []

        @staticmethod
        def dice(y_hat, y, w, axes):
            Metrics.calls.append(tuple(axes))
            return (y_hat * y * w).sum() / ((y_hat + y) * w).sum().clamp_min(1e-6)
        iou = mcc = dice

    logged = {}
    m = iu.UNet(num_classes=2, loss_function=lambda y_hat, y, w, axes: (((y_hat - y) ** 2) * w).mean(dim=axes).sum())
    m.metrics_module = Metrics
    m.log = lambda name, value, **kw: logged.__setitem__(name, float(value))
    batch = (torch.rand(2, 1, 64, 64), torch.rand(2, 2, 64, 64).round(), torch.ones(2, 1, 64, 64))
    m.train()
    opt = m.configure_optimizers()
    before = m.model.segmentation_head[0].weight.detach().clone()
    loss = m.training_step(batch)
    loss.backward()
    opt.step()
    assert not torch.equal(before, m.model.segmentation_head[0].weight)
    assert {"train/Loss", "train/Dice", "train/IoU", "train/MCC"} <= set(logged) and Metrics.calls[0] == (0, 2, 3)
    assert m.validation_step.__code__.co_argcount >= 2


def test_resnet18_encoder_state_dict_and_checkpoint(tmp_path):
    """SURVEY section 8 row f3 (first step): the other BasicBlock ResNet behind the same kernels."""
    import interactive_unet_b200 as iu
    from oracle.smp_unet_resnet34 import RefUNet
    ref = RefUNet(1, 3, "resnet18")
    m = iu.UNet(num_classes=3, encoder_name="resnet18")
    assert list(m.state_dict()) == list(ref.state_dict())
    assert "model.encoder.layer3.1.conv2.weight" in m.state_dict()
    assert "model.encoder.layer3.2.conv1.weight" not in m.state_dict()
    m.load_state_dict(ref.state_dict())
    m.train(), ref.train()
    x = torch.rand(2, 1, 64, 64)
    assert torch.equal(m(x), ref(x))
    path = tmp_path / "r18.ckpt"
    torch.save({"state_dict": ref.state_dict(),
                "hyper_parameters": dict(lr=1e-4, num_channels=1, num_classes=3, architecture="U-Net",
                                         encoder_name="resnet18", pretrained=False)}, path)
    again = iu.UNet.load_from_checkpoint(checkpoint_path=str(path))
    assert all(torch.equal(again.state_dict()[k], v) for k, v in ref.state_dict().items())
    with pytest.raises(RuntimeError):                                # a resnet34 checkpoint does not fit
        iu.UNet(num_classes=3, encoder_name="resnet18").load_state_dict(RefUNet(1, 3).state_dict())
    with pytest.raises(NotImplementedError):
        iu.UNet(encoder_name="resnet50")


def test_load_from_checkpoint_reference_format(tmp_path):
    """Lightning checkpoint layout of `trainer.py:46-49`: state_dict under `model.`, hyper_parameters
    including a pickled `interactive_unet.metrics` loss function."""
    import interactive_unet_b200 as iu
    from interactive_unet_b200.unet import _install_pickle_shims
    from oracle.smp_unet_resnet34 import RefUNet
    _install_pickle_shims()
    loss = sys.modules["interactive_unet.metrics"].mcc_ce_loss
    ref = RefUNet(1, 4)
    path = tmp_path / "model.ckpt"
    torch.save({"state_dict": ref.state_dict(), "epoch": 3,
                "hyper_parameters": dict(lr=2e-4, num_channels=1, num_classes=4, loss_function=loss,
                                         architecture="U-Net", encoder_name="resnet34", pretrained=True)}, path)
    m = iu.UNet.load_from_checkpoint(checkpoint_path=str(path))
    assert m.num_classes == 4 and m.lr == 2e-4
    for k, v in ref.state_dict().items():
        assert torch.equal(m.state_dict()[k], v)


def test_tiling_helpers_match_reference_golden(golden_dir):
    """The drop-in `get_block_coordinates` / `get_padded_block` / `get_shard_coordinates` / `gaussian_3d`
    (predict.py:291-347,362-411) against fixtures recorded from the verbatim reference."""
    from interactive_unet_b200 import predict as P
    g = np.load(os.path.join(golden_dir, "coordinates.npz"))
    n = 0
    while f"case{n}_args" in g:
        a = g[f"case{n}_args"]
        c, p, l = P.get_block_coordinates(np.array(a[:3]), input_size=int(a[3]), overlap=a[4] / 100.0)
        assert np.array_equal(c, g[f"case{n}_clipped"]) and np.array_equal(p, g[f"case{n}_padded"])
        assert np.array_equal(l, g[f"case{n}_local"])
        n += 1
    assert n >= 7
    assert np.array_equal(P.get_shard_coordinates((100, 80, 60), 32), g["shards_100_80_60_32"])
    pb = np.load(os.path.join(golden_dir, "padded_block.npz"))
    for i, box in enumerate(pb["boxes"]):
        assert np.array_equal(P.get_padded_block(pb["volume"], *box), pb[f"out{i}"])
    gw = np.load(os.path.join(golden_dir, "gaussian3d.npz"))
    for size in (8, 16, 32, 48):
        assert np.array_equal(P.gaussian_3d(size), gw[f"w{size}"])


def test_gaussian_window_parameters():
    from interactive_unet_b200 import gaussian_window_1d
    g, gmax, lo = gaussian_window_1d(64)
    assert g.dtype == np.float32 and g.shape == (64,) and g.max() == 1.0 and gmax == 1.0
    assert lo >= 1e-3 - 1e-9


# --------------------------------------------------------------------------- multi-rank host logic (gloo)
class _NumpyEngine:
    """CPU stand-in honouring the Engine's `predict_axis` / `reduce` contracts (layouts of
    include/iunet_b200.h), computing with the oracle port and the exact toy model.  It lets the
    sharding / exchange code in `interactive_unet_b200.distributed` run under gloo without a GPU."""

    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.device = torch.device("cpu")

    def auto_batch(self, h, w, count):
        return min(count, 3)            # small and ragged against the slab thickness: exercises the chunked exchange

    def predict_slices(self, source, offset, count, h, w, strides, out, slice_offset=0, slice_total=None,
                       row_block=None, asynchronous=False, sync=True):
        from oracle import predict_port as pp
        from oracle.make_golden import toy_model_numpy
        c = self.num_classes
        slice_total = count if slice_total is None else slice_total
        row_block = h if row_block is None else row_block
        imgs = torch.as_strided(source.reshape(-1), (count, h, w), tuple(int(v) for v in strides), int(offset)).numpy()
        x = pp.normalise_u8(imgs) if imgs.dtype == np.uint8 else imgs
        p = np.moveaxis(toy_model_numpy(np.ascontiguousarray(x)[:, None], c), 1, -1)
        view = out.view(-1)[:(h // row_block) * slice_total * row_block * w * c].view(h // row_block, slice_total,
                                                                                      row_block, w, c)
        blocks = torch.from_numpy(np.ascontiguousarray(p)).view(count, h // row_block, row_block, w, c)
        view[:, slice_offset:slice_offset + count] = blocks.permute(1, 0, 2, 3, 4)
        return out

    def predict_axis(self, volume, axis, slice_begin=0, slice_count=None, out=None, slice_offset=0, slice_total=None,
                     row_block=None, asynchronous=False):
        vol = torch.as_tensor(volume)
        n, c = vol.shape[0], self.num_classes
        slice_count = n - slice_begin if slice_count is None else slice_count
        slice_total = slice_count if slice_total is None else slice_total
        if out is None:
            out = torch.empty((slice_total, n, n, c), dtype=torch.float32)
        strides = {0: (n * n, n, 1), 1: (n, n * n, 1), 2: (1, n * n, n)}[axis]
        return self.predict_slices(vol.contiguous(), slice_begin * strides[0], slice_count, n, n, strides, out,
                                   slice_offset, slice_total, row_block)

    def reduce(self, probs, order, n, t=None, z0=0, window=None, out_u8=None, out_labels=None, out_mean=None,
               asynchronous=False):
        from oracle import predict_port as pp
        t = n if t is None else t
        c = self.num_classes
        acc = np.zeros((t, n, n, c), np.float32)
        for a in order:
            p = probs[a].numpy().reshape(-1)
            if a == 0:
                acc += p.reshape(t, n, n, c)
            elif a == 1:
                acc += p.reshape(n, t, n, c).transpose(1, 0, 2, 3)          # [y][z][x] -> [z][y][x]
            else:
                acc += p.reshape(n, t, n, c).transpose(1, 2, 0, 3)          # [x][z][y] -> [z][y][x]
        mean = acc / np.float32(len(order))
        if out_mean is not None:
            out_mean.copy_(torch.from_numpy(mean))
        if out_labels is not None:
            out_labels.copy_(torch.from_numpy(pp.labels_from_probs(mean, c).astype(np.uint8)))
        if out_u8 is not None:
            w = pp.gaussian_3d(n)[z0:z0 + t]
            out_u8.copy_(torch.from_numpy(pp.quantise(mean * w[..., None], w)))


def _sharded_worker(rank, world, port, golden_path, result_dir, mode):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from interactive_unet_b200 import distributed as iud
    from interactive_unet_b200 import gaussian_window_1d
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = np.load(golden_path)
        c = int(g["num_classes"])
        vol = torch.from_numpy(g["volume"])
        n = vol.shape[0]
        t = n // world
        # "slab": every rank holds only its z-slab and the strips are exchanged; "volume": replicated input
        src = dict(slab=vol[rank * t:(rank + 1) * t].clone()) if mode == "slab" else dict(volume=vol)
        res = iud.predict_volume_sharded(_NumpyEngine(c), axes=[int(a) for a in g["axes"]],
                                         window=gaussian_window_1d(n), want_mean=True, **src)
        full = iud.gather_slabs(res["u8"])
        np.save(os.path.join(result_dir, f"slab{rank}.npy"), res["u8"].numpy())
        if rank == 0:
            np.save(os.path.join(result_dir, "full.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


def _volume_partition_worker(rank, world, port, result_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from interactive_unet_b200 import distributed as iud
    files = [f"data/image_volumes/v{i}.zarr" for i in range(5)]
    assert iud.volumes_for_rank(files) == files                      # no process group: everything
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        mine = iud.volumes_for_rank(files)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        if rank == 0:
            with open(os.path.join(result_dir, "parts.txt"), "w") as f:
                f.write(repr(everyone))
    finally:
        dist.destroy_process_group()


def test_volume_files_are_partitioned_across_ranks(tmp_path):
    """`predict_volumes` under torchrun: world_size-2 gloo run, every store goes to exactly one rank."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_volume_partition_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    parts = eval(open(tmp_path / "parts.txt").read())
    assert parts == [[f"data/image_volumes/v{i}.zarr" for i in (0, 2, 4)],
                     [f"data/image_volumes/v{i}.zarr" for i in (1, 3)]]


@pytest.mark.parametrize("mode", ["slab", "volume"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_prediction_matches_reference_golden(golden_dir, tmp_path, world, mode):
    """world_size-N gloo run of the z-slab path reproduces the VERBATIM reference `predict_volumes`
    output (golden fixture) bit for bit, i.e. the partition + all-to-all layout is exact."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    golden = os.path.join(golden_dir, "volume_single_s32_c2.npz")
    mp.spawn(_sharded_worker, args=(world, port, golden, str(tmp_path), mode), nprocs=world, join=True)
    want = np.load(golden)["out_u8"]
    assert np.array_equal(np.load(tmp_path / "full.npy"), want)
    t = want.shape[0] // world
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"slab{r}.npy"), want[r * t:(r + 1) * t])


def test_numpy_engine_layout_contract(golden_dir):
    """The stand-in itself follows the documented single-slab layouts (sanity of the test double)."""
    g = np.load(os.path.join(golden_dir, "volume_single_s32_c4.npz"))
    c, n = int(g["num_classes"]), g["volume"].shape[0]
    eng = _NumpyEngine(c)
    from interactive_unet_b200 import gaussian_window_1d
    probs = {a: eng.predict_axis(torch.from_numpy(g["volume"]), a) for a in (0, 1, 2)}
    out = torch.empty((n, n, n, c), dtype=torch.uint8)
    eng.reduce(probs, [0, 1, 2], n, window=gaussian_window_1d(n), out_u8=out)
    assert np.array_equal(out.numpy(), g["out_u8"])

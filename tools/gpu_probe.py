"""Verbose GPU diagnostics for the tensor-core conv kernel and the tail kernels (development aid).

Run on a B200:  python tools/gpu_probe.py [--quick]
Unlike the pytest suite this never stops at the first failure and prints error structure
(per-tap / per-channel) to localise descriptor or layout mistakes in one GPU call.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactive_unet_b200 as iu  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
ACT = torch.float16


def ref_conv(src0, src1, w, b, ksize, stride, residual, relu, up2x):
    x = src0 if src1 is None else torch.cat([src0, src1], dim=3)
    x = x.float().permute(0, 3, 1, 2)
    wq = torch.from_numpy(w).to(dev).to(ACT).float()
    y = F.conv2d(x, wq, torch.from_numpy(b).to(dev), stride=stride, padding=ksize // 2)
    if residual is not None:
        y = y + residual.float().permute(0, 3, 1, 2)
    if relu:
        y = torch.relu(y)
    if up2x:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    return y.permute(0, 2, 3, 1).contiguous()


def conv_case(eng, name, b, h, w, c0, c1, cout, ksize, stride, residual=False, relu=True, up2x=False, seed=0,
              weight_kind="rand"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    src0 = torch.randn(b, h, w, c0, generator=g).to(dev).to(ACT)
    src1 = torch.randn(b, h, w, c1, generator=g).to(dev).to(ACT) if c1 else None
    cin = c0 + c1
    if weight_kind == "delta":
        wt = np.zeros((cout, cin, ksize, ksize), np.float32)
        for o in range(cout):
            wt[o, o % cin, ksize // 2, ksize // 2] = 1.0
    else:
        wt = (torch.randn(cout, cin, ksize, ksize, generator=g) / np.sqrt(cin * ksize * ksize)).numpy()
    bias = (0.1 * torch.randn(cout, generator=g)).numpy() if weight_kind != "delta" else np.zeros(cout, np.float32)
    pad = ksize // 2
    oh, ow = (h + 2 * pad - ksize) // stride + 1, (w + 2 * pad - ksize) // stride + 1
    res = torch.randn(b, oh, ow, cout, generator=g).to(dev).to(ACT) if residual else None
    try:
        t0 = time.time()
        out = eng.conv_test(src0, src1, wt, bias, ksize, stride, residual=res, relu=relu, up2x=up2x)
        torch.cuda.synchronize()
        dt = time.time() - t0
    except Exception as e:  # noqa: BLE001
        print(f"[conv {name}] EXCEPTION: {e}")
        return False
    ref = ref_conv(src0, src1, wt, bias, ksize, stride, res, relu, up2x)
    err = (out.float() - ref).abs()
    tol = 0.02 + 0.01 * ref.abs()
    bad = (err > tol)
    ok = not bool(bad.any())
    print(f"[conv {name}] B{b} {h}x{w} cin {c0}+{c1} -> {cout} k{ksize} s{stride} res={residual} up={up2x}: "
          f"max_err {err.max().item():.4g} mean_err {err.mean().item():.4g} ref_absmax {ref.abs().max().item():.3g} "
          f"bad {bad.float().mean().item():.4f} {'OK' if ok else 'FAIL'} ({dt * 1e3:.1f} ms)")
    if not ok:
        bn = bad.float()
        print("   bad by image  :", [round(v, 3) for v in bn.mean(dim=(1, 2, 3)).tolist()][:16])
        print("   bad by row    :", [round(v, 3) for v in bn.mean(dim=(0, 2, 3)).tolist()][:32])
        print("   bad by col    :", [round(v, 3) for v in bn.mean(dim=(0, 1, 3)).tolist()][:32])
        ch = bn.mean(dim=(0, 1, 2)).tolist()
        print("   bad by channel:", [round(v, 3) for v in ch][:64])
        print("   out[0,0,0,:8] :", out[0, 0, 0, :8].float().tolist())
        print("   ref[0,0,0,:8] :", ref[0, 0, 0, :8].tolist())
        print("   out[0,1,1,:8] :", out[0, 1, 1, :8].float().tolist())
        print("   ref[0,1,1,:8] :", ref[0, 1, 1, :8].tolist())
    return ok


def probe_convs(eng, quick):
    results = []
    cases = [
        ("delta64", dict(b=8, h=16, w=16, c0=64, c1=0, cout=64, ksize=3, stride=1, relu=False, weight_kind="delta")),
        ("1x1_64", dict(b=8, h=16, w=16, c0=64, c1=0, cout=64, ksize=1, stride=1, relu=False)),
        ("3x3_64_64", dict(b=8, h=16, w=16, c0=64, c1=0, cout=64, ksize=3, stride=1)),
        ("3x3_128_128", dict(b=8, h=32, w=32, c0=128, c1=0, cout=128, ksize=3, stride=1)),
        ("3x3_256_256", dict(b=8, h=16, w=16, c0=256, c1=0, cout=256, ksize=3, stride=1)),
        ("cat_64+64_32", dict(b=8, h=32, w=32, c0=64, c1=64, cout=32, ksize=3, stride=1)),
        ("cat_128+64_64", dict(b=8, h=16, w=16, c0=128, c1=64, cout=64, ksize=3, stride=1)),
        ("3x3_32_32", dict(b=8, h=32, w=32, c0=32, c1=0, cout=32, ksize=3, stride=1)),
        ("3x3_32_16", dict(b=8, h=32, w=32, c0=32, c1=0, cout=16, ksize=3, stride=1)),
        ("3x3_16_16", dict(b=8, h=32, w=32, c0=16, c1=0, cout=16, ksize=3, stride=1)),
        ("s2_3x3_64_128", dict(b=8, h=32, w=32, c0=64, c1=0, cout=128, ksize=3, stride=2)),
        ("s2_1x1_64_128", dict(b=8, h=32, w=32, c0=64, c1=0, cout=128, ksize=1, stride=2, relu=False)),
        ("res_relu", dict(b=8, h=16, w=16, c0=64, c1=0, cout=64, ksize=3, stride=1, residual=True)),
        ("up2x", dict(b=8, h=16, w=16, c0=64, c1=0, cout=64, ksize=3, stride=1, up2x=True)),
        ("tiny4x4", dict(b=8, h=4, w=4, c0=512, c1=0, cout=512, ksize=3, stride=1)),
        ("tiny8x8", dict(b=16, h=8, w=8, c0=256, c1=0, cout=256, ksize=3, stride=1)),
        ("odd24x24", dict(b=8, h=24, w=24, c0=64, c1=0, cout=64, ksize=3, stride=1)),
        ("odd3x3", dict(b=8, h=3, w=3, c0=512, c1=0, cout=512, ksize=3, stride=1)),
        ("big64x64", dict(b=8, h=64, w=64, c0=64, c1=0, cout=64, ksize=3, stride=1)),
    ]
    if quick:
        cases = cases[:4]
    for name, kw in cases:
        results.append((name, conv_case(eng, name, **kw)))
    return results


def probe_tail(eng):
    """gather (K1) and reduce (K4) against the numpy port."""
    from oracle import predict_port as pp
    ok_all = True
    n = 64
    rng = np.random.default_rng(5)
    vol = rng.integers(0, 256, (n, n, n), dtype=np.uint8)
    vol_d = torch.from_numpy(vol).to(dev)
    norm = pp.normalise_u8(vol)
    for axis in (0, 1, 2):
        for start, count in ((0, 32), (8, 24), (37, 5)):
            got = eng.gather_slices(vol_d, axis, start, count).cpu().numpy()
            want = pp.slice_batch(norm, axis, start, count)[:, 0]
            ok = np.array_equal(got, want)
            ok_all &= ok
            print(f"[gather u8] axis {axis} start {start} count {count}: {'OK' if ok else 'FAIL'}")
    volf = rng.random((n, n, n), dtype=np.float32)
    for axis in (0, 1, 2):
        got = eng.gather_slices(torch.from_numpy(volf).to(dev), axis, 3, 17).cpu().numpy()
        ok = np.array_equal(got, pp.slice_batch(volf, axis, 3, 17)[:, 0])
        ok_all &= ok
        print(f"[gather f32] axis {axis}: {'OK' if ok else 'FAIL'}")
    # reduce: random per-axis probabilities in slice-major layout
    for c in (2, 4, 3):
        eng.num_classes = c
        p = {a: rng.random((n, n, n, c), dtype=np.float32) for a in (0, 1, 2)}     # [slice][row][col][c]
        acc = np.zeros((n, n, n, c), np.float32)
        order = [0, 1, 2] if c != 3 else [2, 0]
        for a in order:
            pp.scatter_batch(acc, p[a], a, 0)
        mean = acc / np.float32(len(order))
        window = pp.gaussian_3d(n)
        pred, weight = pp.blend_single_block(mean, window)
        want_u8 = pp.quantise(pred, weight)
        want_lab = pp.labels_from_probs(mean, c).astype(np.uint8)
        pd = {a: torch.from_numpy(p[a]).to(dev) for a in order}
        out_u8 = torch.zeros((n, n, n, c), dtype=torch.uint8, device=dev)
        out_lab = torch.zeros((n, n, n), dtype=torch.uint8, device=dev)
        out_mean = torch.zeros((n, n, n, c), dtype=torch.float32, device=dev)
        eng.reduce(pd, order, n, window=iu.gaussian_window_1d(n), out_u8=out_u8, out_labels=out_lab, out_mean=out_mean)
        r = [np.array_equal(out_mean.cpu().numpy(), mean), np.array_equal(out_u8.cpu().numpy(), want_u8),
             np.array_equal(out_lab.cpu().numpy(), want_lab)]
        ok_all &= all(r)
        print(f"[reduce] C={c} order={order}: mean {r[0]} u8 {r[1]} labels {r[2]}"
              f" (u8 mismatches {(out_u8.cpu().numpy() != want_u8).sum()})")
    return ok_all


def probe_network(precision, sizes=(64, 128, 256)):
    """Whole network vs the fp32 oracle (strict fp32 on the GPU), random-init and fitted weights."""
    from oracle import synth
    ok_all = True
    vol, lab = synth.blob_volume(64, 1)
    vol2, _ = synth.blob_volume(256, 2)
    for c, fitted in ((2, False), (2, True), (4, True)):
        ref = synth.make_model(c)
        if fitted:
            ref = synth.fit_decisive(ref, vol, lab % c, steps=100, batch=8, device=dev)
        ref = ref.to(dev).eval()
        model = iu.UNet(num_classes=c)
        model.precision = precision
        model.load_state_dict(ref.state_dict())
        model = model.to(dev).eval()
        for s in sizes:
            x = torch.from_numpy(vol2[:3, :s, :s].astype("float32") / 255.0)[:, None].to(dev)
            with torch.inference_mode():
                want = ref(x)
                t0 = time.time()
                got = model(x)
                torch.cuda.synchronize()
                dt = time.time() - t0
            err = (got - want).abs().max().item()
            agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
            ok = err < 1e-2
            ok_all &= ok
            print(f"[network {precision}] C={c} fitted={fitted} S={s}: max|dp| {err:.4g} argmax agreement {agree:.5f} "
                  f"prob range [{want.min().item():.3f},{want.max().item():.3f}] {'OK' if ok else 'FAIL'} ({dt*1e3:.1f} ms)")
    return ok_all


def probe_volume(precision):
    """predict_volume end to end vs the oracle port driven by the fp32 oracle network on the GPU; then timing."""
    from oracle import predict_port as pp
    from oracle import synth
    n, c = 64, 2
    vol, lab = synth.blob_volume(n, 1)
    ref = synth.fit_decisive(synth.make_model(c), vol, lab, steps=100, batch=8, device=dev).to(dev).eval()
    model = iu.UNet(num_classes=c)
    model.precision = precision
    model.load_state_dict(ref.state_dict())
    model = model.to(dev).eval()

    def fwd(x):
        with torch.inference_mode():
            return ref(torch.from_numpy(x).to(dev)).cpu().numpy()
    want_mean = pp.predict_block(fwd, pp.normalise_u8(vol), c, 16, (0, 1, 2))
    want_u8 = pp.quantise(*pp.blend_single_block(want_mean, pp.gaussian_3d(n)))
    got_mean = iu.predict.predict_block(model, torch.tensor(vol.astype("float32") / 255.0), c, 16, [0, 1, 2])
    got_u8, got_lab = iu.predict.predict_volume_array(model, vol, num_classes=c, return_labels=True)
    d = np.abs(got_mean - want_mean)
    agree = (got_lab == want_mean.argmax(-1)).mean()
    du8 = np.abs(got_u8.astype(int) - want_u8.astype(int))
    print(f"[volume {precision}] N={n}: mean-prob max err {d.max():.4g}  label agreement {agree:.5f}  "
          f"u8 max diff {du8.max()}  u8 exact {(du8 == 0).mean():.4f}")
    # bit-exactness of the tail given the engine's own mean probabilities
    self_u8 = pp.quantise(*pp.blend_single_block(got_mean, pp.gaussian_3d(n)))
    print(f"[volume {precision}] quantise(own mean) bit-exact: {np.array_equal(self_u8, got_u8)}; "
          f"labels from own mean exact: {np.array_equal(got_lab, got_mean.argmax(-1).astype(np.uint8))}")
    for nn in (128, 256):
        v = torch.from_numpy(synth.noise_volume(nn, 3)).to(dev)
        out = torch.empty((nn, nn, nn, c), dtype=torch.uint8, device=dev)
        eng = model.engine()
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.time()
            eng.predict_volume(v, axes=(0, 1, 2), window=iu.gaussian_window_1d(nn), out_u8=out)
            torch.cuda.synchronize()
            dt = time.time() - t0
            print(f"[volume {precision}] N={nn} iter {it}: {dt*1e3:.1f} ms  {nn**3/dt/1e6:.1f} Mvox/s  "
                  f"{nn**3*706.85e3/dt/1e12:.1f} TFLOP/s")


def main():
    global ACT
    quick = "--quick" in sys.argv
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    for precision in ("fp16", "bf16"):
        ACT = torch.float16 if precision == "fp16" else torch.bfloat16
        eng = iu.Engine(0, precision=precision)
        res = probe_convs(eng, quick)
        print(f"conv summary {precision}: all ok = {all(v for _, v in res)}", {k: v for k, v in res if not v})
        if precision == "fp16":
            try:
                print("tail ok:", probe_tail(eng))
            except Exception as e:  # noqa: BLE001
                print("tail EXCEPTION:", repr(e))
        for fn in (probe_network, probe_volume):
            try:
                fn(precision)
            except Exception as e:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                print(fn.__name__, "EXCEPTION:", repr(e))


if __name__ == "__main__":
    main()

"""Benchmark of the full-volume prediction path (BASELINE.json: voxels/s, 3-axis prediction of a
synthetic uint8 volume; N=1 workload = configs[1], 512^3, 2 classes).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete 3-axis prediction of the volume (gather -> U-Net -> softmax -> cross-axis
reduce -> uint8 probabilities + labels).  Prints ONE JSON line on rank 0.

  value        device-resident throughput (volume already in HBM, outputs left in HBM), CUDA events
               on the engine's stream, max over ranks
  e2e          the same prediction through the public drop-in API (`predict.predict_volume_array`)
               with pinned HOST buffers: host->device copy of the volume and device->host copy of the
               uint8 probabilities + labels inside the timed region
  roofline     the tcgen05 conv kernels (tensor bound): algorithmic conv FLOPs / their summed device
               time (CUDA events around every launch, separate profiled pass); HBM-bound kernels
               (gather K1, reduce K4) are reported in `roofline_hbm`
  cpu_baseline the oracle (fp32 restatement of the reference network + port of predict.py) on this
               box's host cores, on a bounded sample of the same workload
`--impl reference` times that CPU path as its own arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_VOXEL_AXIS = {2: 235.62e3, 4: 236.19e3}          # BASELINE.md section 3 (dense conv FLOPs per pixel per axis)
STEM_FLOP_PER_VOXEL_AXIS = 2 * 49 * 64 / 4.0               # the 7x7/s2 stem's share of the figure above (its own kernel class)
# N = 1 is BASELINE.json configs[1] (512^3 on one B200); N = 2 / 4 / 8 run configs[2]'s 1024^3 volume z-slab sharded
# (512 / 256 / 128 slices of 1024^2 per GPU and axis).  The per-voxel work is the same at every N, so value(N) / N is
# comparable across the row; the edges between 512 and 1024 that would keep the voxel count per GPU exactly constant
# (640, 800) have feature maps that tile badly (20x20, 25x25) and would measure that instead of the sharding.
WEAK_EDGES = {1: 512, 2: 1024, 4: 1024, 8: 1024}
AXES = (0, 1, 2)
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_conv_dram.json")   # written by tools/ncu_dram.py from an ncu capture


ENCODER = "resnet34"                                       # --encoder
ENCODER_FLOP_DELTA = {"resnet34": 0.0, "resnet18": -73.728e3}    # resnet18 has 8 fewer BasicBlocks (first step)


def flops_per_voxel(classes, n_axes=3, encoder="resnet34"):
    per_axis = FLOP_PER_VOXEL_AXIS.get(classes, 235.62e3 + (classes - 2) * 0.285e3) + ENCODER_FLOP_DELTA[encoder]
    return per_axis * n_axes


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tc_tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="MEASURED_PEAKS.json (hbm_gbs, bf16_tflops_sustained)")
    return dict(hbm_gbs=6650.0, tc_tflops=1590.0, source="fallback of B200_PROFILING.md (6.65 TB/s, 1.59 PFLOP/s)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                power.append(float(f[2]))                  # board power: the step sits on the cap
            except ValueError:
                pass
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        if power:
            out["power_w"] = float(np.median(power))
        return out


# ------------------------------------------------------------------------------------------- synthetic inputs
def synthetic_model(classes, encoder, seed=1234, calib_size=64):
    """Random-init weights of the named architecture in the product's own parameter container (no checkpoint can be
    downloaded here): BN affine parameters randomised, BN statistics calibrated on noise and then jittered, so that
    BN folding is exercised and activations stay O(1) through the 47 layers.  The CPU arm
    loads the same state_dict into the oracle network."""
    import interactive_unet_b200 as iu
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    model = iu.UNet(num_classes=classes, encoder_name=encoder)
    net = model.model
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        for m in bns:
            m.weight.copy_(0.5 + torch.rand(m.num_features, generator=gen))
            m.bias.copy_(-0.2 + 0.4 * torch.rand(m.num_features, generator=gen))
            m.reset_running_stats()
            m.momentum = None                       # cumulative average over the calibration batches
        net.train()
        for _ in range(2):
            net(torch.rand(4, 1, calib_size, calib_size, generator=gen))
        net.eval()
        for m in bns:
            m.momentum = 0.1
            m.running_mean += 0.1 * m.running_var.sqrt() * torch.randn(m.num_features, generator=gen)
            m.running_var *= 0.75 + 0.5 * torch.rand(m.num_features, generator=gen)
        net.segmentation_head[0].bias.copy_(0.1 * torch.randn(classes, generator=gen))
    return model.eval()


def noise_volume(n, seed):
    """Uniform uint8 noise [n,n,n] (the value distribution does not affect timing)."""
    return np.random.default_rng(seed).integers(0, 256, (n, n, n), dtype=np.uint8)


def noise_slab(n, z0, t, seed):
    """Planes [z0, z0 + t) of a uint8 noise volume that is the same whatever the number of ranks (one generator per
    plane): every rank makes only its own slab."""
    out = np.empty((t, n, n), dtype=np.uint8)
    for i in range(t):
        out[i] = np.random.default_rng([seed, z0 + i]).integers(0, 256, (n, n), dtype=np.uint8)
    return out


def conv_dram_traffic(edge, world):
    """`roofline.traffic`: DRAM bytes per conv launch from the committed ncu capture (tools/ncu_dram.py), or None when
    no capture of this workload is on file."""
    try:
        rec = json.load(open(NCU_TRAFFIC_FILE))
    except (OSError, ValueError):
        return None, None
    if rec.get("edge") != edge or world != 1:
        return None, rec.get("source")
    return rec.get("dram_bytes_per_launch"), rec.get("source")


def torch_gpu_baseline(model, dev, edge, classes, slices=32, repeats=3):
    """The same network in PyTorch eager / cuDNN on this GPU (network only: no gather, accumulate or tail), as
    voxels/s-equivalent of a 3-axis prediction, timed with CUDA events on torch's stream after a warm-up: context for
    `value`.  Two ways: as the reference runs it (fp32 NCHW, TF32 convolutions allowed -- torch's default, which the
    reference does not change, predict.py:124 only sets the matmul precision) and tuned (fp16, channels_last)."""
    import copy
    out = {"unit": "voxels/s", "sample": f"{slices} slices of {edge}x{edge}, best of {repeats}",
           "note": "network only (no gather / reduce / tail); BatchNorm in eval mode"}

    def timed(net, x):
        with torch.inference_mode():
            for _ in range(2):
                torch.softmax(net(x), 1)
            torch.cuda.synchronize(dev)
            best = None
            for _ in range(repeats):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.softmax(net(x), 1)
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
        return slices * edge * edge / 3.0 / (best * 1e-3)
    try:
        net = copy.deepcopy(model.model).to(dev).eval()
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        out["as_reference_fp32_tf32_nchw"] = timed(net, torch.rand(slices, 1, edge, edge, device=dev))
        torch.backends.cudnn.allow_tf32 = tf32
        half = net.half().to(memory_format=torch.channels_last)
        x = torch.rand(slices, 1, edge, edge, device=dev).half().contiguous(memory_format=torch.channels_last)
        out["tuned_fp16_channels_last"] = timed(half, x)
        out["value"] = out["tuned_fp16_channels_last"]
        out["kind"] = "PyTorch eager + cuDNN, fp16 channels_last (the faster of the two)"
        del net, half, x
    except Exception as exc:            # context only: never fail the bench over it
        out["unavailable"] = f"{type(exc).__name__}: {exc}"[:200]
    torch.cuda.empty_cache()
    return out


def cpu_config0(threads, repeats=3):
    """BASELINE.json configs[0], as specified: the reference path (port of predict.py:79-112 + the restated fp32
    network) on a synthetic 128^3 uint8 volume, 2 classes, axes=[0], batch 128, on the host cores: one warm-up, best of
    `repeats`."""
    from oracle import predict_port as pp
    from oracle.smp_unet_resnet34 import RefUNet
    torch.set_num_threads(threads)
    net = RefUNet(1, 2, "resnet34")
    net.load_state_dict(synthetic_model(2, "resnet34").state_dict())
    net.eval()
    vol = np.random.default_rng(0).integers(0, 256, (128, 128, 128), dtype=np.uint8)

    def fwd(x):
        with torch.inference_mode():
            return net(torch.from_numpy(x)).numpy()
    best = None
    for i in range(repeats + 1):
        t0 = time.perf_counter()
        pp.predict_block(fwd, pp.normalise_u8(vol), 2, 128, (0,))
        dt = time.perf_counter() - t0
        if i > 0:
            best = dt if best is None else min(best, dt)
    return {"workload": "configs[0]: single-axis prediction of a synthetic 128^3 uint8 volume, 2 classes, CPU",
            "value": 128 ** 3 / best, "unit": "voxels/s", "seconds": best, "cores": threads, "batch_size": 128,
            "kind": "port"}


# ------------------------------------------------------------------------------------------- CPU arm
_CPU_SETUP = {}


def cpu_reference_sample(edge, classes, slices_per_axis, threads, repeats=1):
    """Oracle on host cores: `slices_per_axis` slices of edge^2 along each of the 3 axes of the bench
    volume through the fp32 network + the port's scatter / average / quantise.  Returns voxels/s where one
    voxel = one 3-axis prediction (3 slice-pixels), and the seconds of the best repeat."""
    from oracle import predict_port as pp
    from oracle.smp_unet_resnet34 import RefUNet
    torch.set_num_threads(threads)
    key = (edge, classes)
    if key not in _CPU_SETUP:                       # weights / volume are set-up, not part of the timed sample
        oracle_net = RefUNet(1, classes, ENCODER)
        oracle_net.load_state_dict(synthetic_model(classes, ENCODER).state_dict())
        _CPU_SETUP[key] = (oracle_net.eval(), noise_volume(edge, 1))
    model, vol = _CPU_SETUP[key]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        acc = np.zeros((slices_per_axis, edge, edge, classes), np.float32)
        with torch.inference_mode():
            for axis in AXES:
                x = pp.slice_batch(pp.normalise_u8(np.moveaxis(vol, axis, 0)[:slices_per_axis]), 0, 0, slices_per_axis)
                p = model(torch.from_numpy(x)).numpy()
                acc += np.moveaxis(p, 1, -1)
        acc /= np.float32(len(AXES))
        w = pp.gaussian_3d(edge)[:slices_per_axis] if edge <= 256 else np.ones((slices_per_axis, edge, edge), np.float32)
        pp.quantise(acc * w[..., None], w)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    voxels = slices_per_axis * edge * edge
    return voxels / best, best


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    edge = WEAK_EDGES.get(args.gpus, 512) if args.edge is None else args.edge
    threads = os.cpu_count() or 1
    spa = args.cpu_slices
    for _ in range(args.warmup):
        cpu_reference_sample(edge, args.classes, 1, threads)
    t0 = time.perf_counter()
    vals = [cpu_reference_sample(edge, args.classes, spa, threads)[0] for _ in range(args.steps)]
    total = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = (f"{spa} slices of {edge}x{edge} per axis x 3 axes per step (of {edge} per axis): numpy port of predict.py "
              f"+ restated fp32 network (oracle/), torch CPU, {threads} threads; ms_per_step is the SAMPLE's time")
    line = {
        "impl": "reference", "metric": "voxels/sec, full 3-axis volume prediction", "value": value, "unit": "voxels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"3-axis prediction of a synthetic {edge}^3 uint8 volume, {args.classes} classes",
                   "edge": edge, "classes": args.classes, "axes": list(AXES),
                   "network": f"smp.Unet({ENCODER}) restated (oracle/), reference predict.py arithmetic (port)"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_config0:
        line["config0"] = cpu_config0(threads)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import interactive_unet_b200 as iu
    from interactive_unet_b200 import distributed as iud

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    classes = args.classes
    edge = WEAK_EDGES.get(world, 512) if args.edge is None else args.edge
    if edge % world or edge % 32:
        raise SystemExit(f"edge {edge} must be divisible by 32 and by the number of GPUs {world}")
    t_slab = edge // world
    model = synthetic_model(classes, ENCODER)                         # random-init weights of the named architecture
    model.precision = args.precision
    model = model.to(dev).eval()
    eng = model.engine()
    window = iu.gaussian_window_1d(edge)

    # one GPU: the whole volume; several: every rank makes, holds and uploads ONLY its z-slab (the strips it needs
    # along the other two axes are exchanged between the GPUs, distributed.py)
    if world == 1:
        vol_host = torch.from_numpy(noise_volume(edge, 1)).pin_memory()
    else:
        vol_host = torch.from_numpy(noise_slab(edge, rank * t_slab, t_slab, 1)).pin_memory()
    vol_dev = vol_host.to(dev)
    stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident step
    if world == 1:
        out_u8 = torch.empty((edge, edge, edge, classes), dtype=torch.uint8, device=dev)
        out_lab = torch.empty((edge, edge, edge), dtype=torch.uint8, device=dev)

        def step():
            eng.predict_volume(vol_dev, axes=AXES, window=window, out_u8=out_u8, out_labels=out_lab)
    else:
        def step():
            return iud.predict_volume_sharded(eng, slab=vol_dev, axes=AXES, window=window)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.launch_count()
    with ClockSampler(local_rank) as clocks:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
        dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0

    # ---- end-to-end step: pinned host volume in, pinned host uint8 probabilities + labels out
    if world == 1:
        h_u8 = torch.empty((edge, edge, edge, classes), dtype=torch.uint8).pin_memory()
        h_lab = torch.empty((edge, edge, edge), dtype=torch.uint8).pin_memory()

        def e2e_step():
            iu.predict.predict_volume_array(model, vol_host.numpy(), num_classes=classes, axes=list(AXES),
                                            return_labels=True, out=h_u8.numpy(), out_labels=h_lab.numpy())
        h2d, d2h = edge ** 3, edge ** 3 * (classes + 1)
    else:
        h_u8 = torch.empty((t_slab, edge, edge, classes), dtype=torch.uint8).pin_memory()
        h_lab = torch.empty((t_slab, edge, edge), dtype=torch.uint8).pin_memory()

        def e2e_step():
            iud.predict_slab_from_host(eng, vol_host, axes=AXES, window=window, out_u8=h_u8, out_labels=h_lab)
        h2d, d2h = edge ** 3, edge ** 3 * (classes + 1)               # summed over the ranks: each moves its slab only
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- per-kernel-class device time (separate, profiled pass; not part of `value`)
    eng.profile(True)
    eng.profile_read(reset=True)
    prof_steps = max(1, min(args.steps, 3))
    for _ in range(prof_steps):
        step()
    prof = eng.profile_read(reset=True)
    eng.profile(False)

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(times[0]), float(times[1])
    if rank != 0:
        return
    voxels = float(edge) ** 3
    peaks = measured_peaks()
    value = voxels * args.steps / (dev_ms * 1e-3)
    conv_ms, conv_n = prof["conv"]
    # the conv class is every tcgen05 3x3 / 1x1 conv and the head; the stem is its own kernel class, so its FLOPs are
    # left out of this numerator (`network_tflops` below has every FLOP over every network kernel's time)
    conv_fpv = flops_per_voxel(classes, encoder=ENCODER) - 3 * STEM_FLOP_PER_VOXEL_AXIS
    conv_flops = conv_fpv * voxels / world * prof_steps                                        # this rank's share
    conv_tflops = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    net_ms = conv_ms + prof["stem"][0] + prof["pool"][0]
    net_tflops = flops_per_voxel(classes, encoder=ENCODER) * voxels / world * prof_steps / (net_ms * 1e-3) / 1e12 \
        if net_ms > 0 else 0.0
    traffic, traffic_source = conv_dram_traffic(edge, world)
    gather_ms, gather_n = prof["gather"]
    reduce_ms, reduce_n = prof["reduce"]
    share = voxels / world * prof_steps
    gather_gbs = share * 3 * (1 + 4) / (gather_ms * 1e-3) / 1e9 if gather_ms > 0 else 0.0
    reduce_bytes_per_voxel = 3 * classes * 4 + classes + 1
    reduce_gbs = share * reduce_bytes_per_voxel / (reduce_ms * 1e-3) / 1e9 if reduce_ms > 0 else 0.0
    total_prof_ms = sum(v[0] for v in prof.values())
    line = {
        "metric": "voxels/sec, full 3-axis volume prediction", "value": value, "unit": "voxels/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "impl": "ours",
        "config": {"workload": f"3-axis prediction of a synthetic {edge}^3 uint8 volume, {classes} classes, "
                               f"{'1 B200' if world == 1 else f'z-slab sharded across {world} B200'}",
                   "edge": edge, "classes": classes, "axes": list(AXES), "network": f"smp.Unet({ENCODER}), random init",
                   "storage": f"{args.precision} activations/weights, fp32 accumulate (TMEM), fp32 tail",
                   "l2": "inputs_exceed_l2 (per-step working set of several GB >> 126 MB L2; no explicit flush)",
                   "slices_per_gpu_per_axis": t_slab,
                   "input": "whole volume on the one GPU" if world == 1 else
                            "each rank holds / uploads its z-slab; uint8 strips and fp32 partial probabilities of the "
                            "off-slab axes exchanged with NCCL all-to-all"},
        "e2e": {"value": voxels * args.steps / (e2e_ms * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                "api": "interactive_unet_b200.predict.predict_volume_array (pinned host buffers)" if world == 1
                       else "interactive_unet_b200.distributed.predict_slab_from_host (pinned host buffers)"},
        "gpu_launches": int(launches),
        "algorithmic_tflops": flops_per_voxel(classes, encoder=ENCODER) * value / 1e12,
        "roofline": {"kernel": "conv_tc_kernel (tcgen05 implicit-GEMM, all conv layers + head)", "bound": "tensor",
                     "achieved": conv_tflops, "peak": peaks["tc_tflops"], "unit": "TFLOP/s",
                     "frac": conv_tflops / peaks["tc_tflops"],
                     # DRAM bytes per conv launch (read + write) from the committed `ncu --set full` capture of this
                     # workload (profiles/ncu_conv_dram.json, tools/ncu_dram.py); null when none is on file
                     "traffic": traffic, "traffic_source": traffic_source,
                     "peak_source": peaks["source"],
                     "flops_per_voxel": conv_fpv, "launches": int(conv_n),
                     "kernel_ms_per_step": conv_ms / prof_steps,
                     "share_of_step": conv_ms / total_prof_ms if total_prof_ms else None,
                     "network_tflops": net_tflops, "network_frac": net_tflops / peaks["tc_tflops"],
                     "network_note": "all conv FLOPs incl. the stem over conv + stem + max-pool kernel time"},
        "roofline_hbm": {
            "gather": {"bound": "hbm", "achieved": gather_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": gather_gbs / peaks["hbm_gbs"], "bytes_per_voxel_per_axis": 5,
                       "kernel_ms_per_step": gather_ms / prof_steps, "launches": int(gather_n)},
            "reduce": {"bound": "hbm", "achieved": reduce_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": reduce_gbs / peaks["hbm_gbs"], "bytes_per_voxel": reduce_bytes_per_voxel,
                       "kernel_ms_per_step": reduce_ms / prof_steps, "launches": int(reduce_n)}},
        "kernel_ms_per_step": {k: v[0] / prof_steps for k, v in prof.items()},
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, secs = cpu_reference_sample(edge, classes, args.cpu_slices, threads)
        line["cpu_baseline"] = {"value": v, "unit": "voxels/s", "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_slices} slices of {edge}x{edge} per axis x 3 axes "
                                          f"({secs:.1f} s): numpy port of predict.py (not the verbatim module) + "
                                          f"restated fp32 network (oracle/), torch CPU"}
        if not args.no_config0:
            line["cpu_baseline"]["config0"] = cpu_config0(threads)
    if world == 1 and not args.no_gpu_baseline:
        line["gpu_baseline"] = torch_gpu_baseline(model, dev, edge, classes)
    print(json.dumps(line), flush=True)


def main():
    global ENCODER
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--edge", type=int, default=None, help="volume edge (default: 512 at 1 GPU, weak-scaled above)")
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--encoder", default="resnet34", choices=sorted(ENCODER_FLOP_DELTA),
                    help="BASELINE.json's configuration is resnet34; resnet18 is the other supported BasicBlock encoder")
    ap.add_argument("--precision", default=os.environ.get("IU_PRECISION", "fp16"), choices=["fp16", "bf16"])
    ap.add_argument("--cpu-slices", type=int, default=32, help="slices per axis in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the PyTorch/cuDNN-on-this-GPU context number")
    ap.add_argument("--no-config0", action="store_true", help="skip timing BASELINE configs[0] (128^3, 1 axis, CPU)")
    args = ap.parse_args()
    ENCODER = args.encoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # exactly ONE line may reach stdout (the driver parses it): libraries that print there (NCCL's version banner,
    # ...) are sent to stderr for the whole run, and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved_stdout, "w")
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

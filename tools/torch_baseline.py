"""PyTorch / cuDNN on the same B200, for context (SURVEY.md section 8d "GPU baseline"): the oracle's restated network
run the way the reference would run it on a GPU -- eager fp32 with TF32 convolutions, NCHW -- and a tuned variant
(channels_last, bf16 autocast, BatchNorm folded by eval mode), on batches of 512x512 slices.  Reports slices/s and
the equivalent voxels/s of a 3-axis prediction (3 slice-pixels per voxel).  Not part of bench.py's contract.

    python tools/torch_baseline.py [--batch 32] [--iters 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--size", type=int, default=512)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    model = synth.make_model(2).to(dev).eval()
    x = torch.rand(args.batch, 1, args.size, args.size, device=dev)
    out = {"batch": args.batch, "size": args.size, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    px = args.batch * args.size * args.size

    def report(name, ms):
        out[name] = {"ms_per_batch": ms, "slices_per_s": args.batch / (ms * 1e-3), "voxels_per_s_3axis": px / 3 / (ms * 1e-3)}

    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    with torch.inference_mode():
        report("eager_fp32_tf32_nchw", timed(lambda: model(x), args.iters))
        m2 = model.to(memory_format=torch.channels_last)
        x2 = x.to(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            report("eager_bf16_autocast_channels_last", timed(lambda: m2(x2), args.iters))
        m3 = model.half().to(memory_format=torch.channels_last)
        x3 = x2.half()
        report("eager_fp16_channels_last", timed(lambda: m3(x3), args.iters))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

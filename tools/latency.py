"""Latency of the interactive path (BASELINE.json config 4): one `UNet.forward` + argmax on a batch of 256x256 slices,
weights resident, p50 / p99 over many iterations; plus one tiled-mode volume (config "what the GUI really runs").

    python tools/latency.py [--iters 1000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactive_unet_b200 as iu  # noqa: E402
from oracle import synth  # noqa: E402  (seeded synthetic weights / volumes only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--size", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    model = iu.UNet(num_classes=2)
    model.load_state_dict(synth.make_model(2).state_dict())
    model = model.to(dev).eval()
    out = {}
    for batch in (1, 64):
        x = torch.rand(batch, 1, args.size, args.size, device=dev)
        iters = args.iters if batch == 1 else max(50, args.iters // 10)
        for _ in range(20):
            model(x).argmax(1)
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter()
            lab = model(x).argmax(1)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        out[f"forward_argmax_b{batch}_{args.size}"] = {"p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)),
                                                      "mean_ms": float(ts.mean()), "iters": iters,
                                                      "slices_per_s": float(batch / (np.percentile(ts, 50) * 1e-3))}
        del lab
    # tiled mode: 512^3 volume in 256^3 blocks with 25 % overlap (27 blocks, 3.4x the voxels of the volume)
    vol = torch.from_numpy(synth.noise_volume(512, 1)).to(dev)
    iu.predict.predict_volume_array(model, vol[:256, :256, :256].contiguous(), input_size=256, num_classes=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    iu.predict.predict_volume_array(model, vol, input_size=256, num_classes=2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["tiled_512_in_256_blocks"] = {"seconds": dt, "blocks": 27, "volume_voxels_per_s": 512 ** 3 / dt,
                                      "block_voxels_per_s": 27 * 256 ** 3 / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

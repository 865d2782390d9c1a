"""One small 3-axis volume prediction (development aid for ncu captures of the HBM-bound kernels K1 / K4 / pool / stem).

    python tools/profile_volume.py [--edge 256]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactive_unet_b200 as iu  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    model = iu.UNet(num_classes=2)
    model.load_state_dict(synth.make_model(2).state_dict())
    model = model.to(dev).eval()
    vol = torch.from_numpy(synth.noise_volume(args.edge, 1)).to(dev)
    for _ in range(2):
        u8, lab = iu.predict.predict_volume_array(model, vol, num_classes=2, return_labels=True)
    torch.cuda.synchronize()
    print("checksum", int(u8.sum()), int(lab.sum()))


if __name__ == "__main__":
    main()

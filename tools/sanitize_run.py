"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once.

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactive_unet_b200 as iu  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    ref = synth.make_model(2)
    model = iu.UNet(num_classes=2)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev).eval()
    vol = synth.blob_volume(64, 1)[0]
    u8, lab = iu.predict.predict_volume_array(model, vol, num_classes=2, return_labels=True)           # 64^3, 3 axes
    tiled = iu.predict.predict_volume_array(model, vol[:48, :64, :40].copy(), input_size=32, num_classes=2)  # tiled mode
    x = torch.rand(1, 1, 128, 512, device=dev)                                                          # fused tail, row kernels
    with torch.inference_mode():
        y = model(x)
    print("ok", int(u8.sum()), int(lab.sum()), int(tiled.sum()), float(y.sum()))


if __name__ == "__main__":
    main()

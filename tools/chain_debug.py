"""Where does the fused decoder tail differ from the three separate launches?  (development aid)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactive_unet_b200 as iu  # noqa: E402
from oracle import synth  # noqa: E402


def run(chain, ref, x, c):
    os.environ["IU_CONV_CHAIN"] = chain
    model = iu.UNet(num_classes=c)
    model.load_state_dict(ref.state_dict())
    model = model.to(x.device).eval()
    with torch.inference_mode():
        return model(x).clone()


def main():
    dev = torch.device("cuda:0")
    c, h, w = 2, 512, 512
    ref = synth.make_model(c)
    x = torch.rand(2, 1, h, w, generator=torch.Generator().manual_seed(1)).to(dev)
    a, b = run("1", ref, x, c), run("0", ref, x, c)
    d = (a - b).abs().amax(1)                       # [B, H, W]
    bad = d > 2e-3
    print("max diff", float(d.max()), "bad fraction", float(bad.float().mean()))
    print("bad fraction per image", [float(bad[i].float().mean()) for i in range(bad.shape[0])])
    rows = bad.float().mean((0, 2)).cpu().numpy()
    cols = bad.float().mean((0, 1)).cpu().numpy()
    print("bad by row mod 8 ", np.round([rows[k::8].mean() for k in range(8)], 3))
    print("bad rows 0..23   ", np.round(rows[:24], 2))
    print("bad rows last 16 ", np.round(rows[-16:], 2))
    print("bad by col mod 124", np.round([cols[k::124].mean() for k in range(0, 124, 8)], 3))
    print("bad cols 0..15   ", np.round(cols[:16], 2))
    print("bad cols 116..140", np.round(cols[116:140], 2))
    print("bad per strip    ", np.round([cols[s * 124:(s + 1) * 124].mean() for s in range(5)], 3))
    print("mean diff by row block of 8 (first 10)", np.round([float(d[:, 8 * k:8 * k + 8].mean()) for k in range(10)], 4))
    print("sample a", a[0, 0, 100, 100:108].cpu().numpy())
    print("sample b", b[0, 0, 100, 100:108].cpu().numpy())
    print("nan in a", bool(torch.isnan(a).any()))


if __name__ == "__main__":
    main()

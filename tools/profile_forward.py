"""One network forward on a batch of synthetic slices (development aid for ncu captures).

    python tools/profile_forward.py [--size 512] [--batch 32] [--iters 2]

Each forward is exactly one pass of every layer kernel (stem, pool, 43 tensor-core convs), so
`ncu -k regex:conv_ -s <launches of iter 1> -c <launches of iter 2>` captures one warm pass.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactive_unet_b200 as iu  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--classes", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ref = synth.make_model(args.classes)
    model = iu.UNet(num_classes=args.classes)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev).eval()
    model.engine().set_max_batch(args.batch)
    x = torch.rand(args.batch, 1, args.size, args.size, device=dev)
    for i in range(args.iters):
        torch.cuda.synchronize()
        t0 = time.time()
        y = model(x)
        torch.cuda.synchronize()
        print(f"iter {i}: {1e3 * (time.time() - t0):.2f} ms, launches so far {model.engine().launch_count()}")
    print("checksum", float(y.sum()))
    if os.environ.get("IU_CONV_DEBUG"):
        eng = model.engine()
        eng.debug_counters(reset=True)
        model(x)
        c = eng.debug_counters(reset=True).astype(float)
        print("layer ctas | MMA: total  w_acc  w_A  w_B  issue | gather: total w_empty issue w_land | epi: wait body | CTA life  (kcycles per CTA)")
        for i, r in enumerate(c):
            n = r[10]
            if n == 0:
                continue
            k = 1e-3 / n
            issue = r[3] - r[0] - r[1] - r[2]
            print(f"{i:3d} {int(n):4d} | {r[3]*k:8.1f} {r[0]*k:7.1f} {r[1]*k:7.1f} {r[2]*k:7.1f} {issue*k:7.1f} | "
                  f"{r[7]*k:8.1f} {r[4]*k:7.1f} {r[5]*k:7.1f} {r[6]*k:7.1f} | {r[8]*k:7.1f} {r[9]*k:7.1f} | {r[11]*k:8.1f}  rest {r[12]*k:6.1f}")


if __name__ == "__main__":
    main()

"""Where 16-bit storage costs accuracy, layer by layer (CPU; test infrastructure + oracle only, no GPU).

BASELINE.json's north_star names bf16 operands AND a 1e-2 probability tolerance.  This script runs the oracle network
on fitted weights through the engine's arithmetic contract (`tests/bf16_emulation.py`: BatchNorm folded, weights and
inter-layer activations rounded to the storage format, fp32 accumulation) for several storage choices and prints, per
network depth, the error of the activations against the fp32 network, and the final probability error.

    python tools/precision_table.py > profiles/r02_precision_by_depth.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bf16_emulation import emulate_logits  # noqa: E402
from oracle import synth  # noqa: E402

MODES = [
    ("fp16 storage (engine default)", dict(act="fp16")),
    ("bf16 storage (IU_PRECISION=bf16)", dict(act="bf16")),
    ("bf16 weights, fp16 activations", dict(act="fp16", weights="bf16")),
    ("bf16 operands, fp32 shortcut stream", dict(act="bf16", residual_stream="fp32")),
    ("bf16 operands, fp16 shortcut stream", dict(act="bf16", residual_stream="fp16")),
]
TAPS = ["stem", "layer1.2", "layer2.3", "layer3.5", "layer4.2", "dec0", "dec1", "dec2", "dec3", "dec4", "logits"]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    vol, lab = synth.blob_volume(64, 1)
    print("fitting the oracle network for 100 AdamW steps (decisive softmax, as in tests/test_gpu_parity.py) ...", file=sys.stderr)
    ref = synth.fit_decisive(synth.make_model(2), vol, lab % 2, steps=100, batch=8).eval()
    big, _ = synth.blob_volume(128, 5)
    x = torch.from_numpy(np.tile(big[:2], (1, 2, 2)).astype(np.float32) / 255.0)[:, None]        # two 256^2 slices
    exact = {}
    logits32 = emulate_logits(ref, x, quantise=False, taps=exact)
    p32 = torch.softmax(logits32, 1)
    print("Activation error against the fp32 network by depth: max|a - a32| / max|a32| per tap; two 256x256 slices, weights")
    print("fitted for 100 steps.  Last columns: max-abs logit and probability error (the gate is 1e-2 on probabilities).\n")
    print(f"{'storage':38s} " + " ".join(f"{t:>9s}" for t in TAPS) + f" {'max|dp|':>9s} {'argmax=':>8s}")
    for name, kw in MODES:
        taps = {}
        logits = emulate_logits(ref, x, quantise=True, taps=taps, **kw)
        p = torch.softmax(logits, 1)
        cells = []
        for t in TAPS:
            a, b = taps[t], exact[t]
            cells.append(float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)))
        dp = float((p - p32).abs().max())
        agree = float((p.argmax(1) == p32.argmax(1)).float().mean())
        print(f"{name:38s} " + " ".join(f"{c:9.2e}" for c in cells) + f" {dp:9.2e} {agree:8.5f}")
    print("\nReading: the error is set by the 8-bit significand of EVERY 16-bit rounding (weights and activations alike);")
    print("keeping only the shortcut stream wider, or only the activations, removes less than half of it, so no bf16-operand")
    print("variant reaches the 1e-2 gate on these weights, while fp16 storage (same tcgen05 kind::f16 rate, same bytes)")
    print("sits an order of magnitude inside it.  fp16's range (65504) is covered by the saturating pack in every epilogue")
    print("and by tests/test_gpu_parity.py::test_fp16_range_stress.")


if __name__ == "__main__":
    main()

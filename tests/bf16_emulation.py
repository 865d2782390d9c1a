"""CPU emulation of the engine's numerics (test infrastructure, not product code).

Runs the oracle network's weights through the same arithmetic contract as the CUDA engine --
BatchNorm folded into the conv weights, weights and inter-layer activations rounded to bf16, fp32
accumulation, fused residual / downsample accumulation, fp32 stem -- using stock PyTorch CPU ops.
It predicts how far a *correct* bf16 engine is expected to sit from the fp32 oracle, and gives a
tight reference for debugging the CUDA path layer by layer.
"""
import torch
import torch.nn.functional as F


def _q(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _q16(t):
    return t.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)


QUANTISERS = {"bf16": _q, "fp16": _q16, "fp32": lambda t: t}


def _fold(conv_w, bn):
    s = (bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps))
    w = (conv_w.double() * s[:, None, None, None]).float()
    b = (bn.bias.double() - bn.running_mean.double() * s).float()
    return w, b


@torch.no_grad()
def emulate_logits(ref_model, x, quantise=True, taps=None, act="bf16", weights=None, residual_stream=None):
    """`ref_model`: oracle RefUNet (eval).  x: fp32 [B,1,H,W].  Returns fp32 logits [B,C,H,W].
    `taps` (optional dict) receives intermediate activations by name.
    `act` / `weights`: storage format of activations / packed weights ("bf16", "fp16", "fp32"; weights default to
    `act`).  `residual_stream`: format the ResNet blocks' outputs are KEPT in for the shortcut addition while the GEMMs
    still read them rounded to `act` (None = same as `act`, which is what the engine does)."""
    qa = QUANTISERS[act] if quantise else (lambda t: t)
    qw = QUANTISERS[weights or act] if quantise else (lambda t: t)
    qr = QUANTISERS[residual_stream] if (quantise and residual_stream) else None
    net = ref_model.model
    e = net.encoder

    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    w, b = _fold(e.conv1.weight, e.bn1)
    f1 = rec("stem", qa(F.relu(F.conv2d(x, w, b, stride=2, padding=3))))       # stem runs in fp32, output 16-bit
    y = rec("pool", F.max_pool2d(f1, 3, 2, 1))
    keep = y                                                                   # the shortcut's copy of the stream
    feats = [f1]
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(e, f"layer{li}")):
            w1, b1 = _fold(blk.conv1.weight, blk.bn1)
            t = qa(F.relu(F.conv2d(y, qw(w1), b1, stride=blk.conv1.stride, padding=1)))
            w2, b2 = _fold(blk.conv2.weight, blk.bn2)
            acc = F.conv2d(t, qw(w2), b2, padding=1)
            if blk.downsample is not None:
                wd, bd = _fold(blk.downsample[0].weight, blk.downsample[1])
                acc = acc + F.conv2d(y, qw(wd), bd, stride=2)
            else:
                acc = acc + (keep if qr is not None else y)
            out = F.relu(acc)
            keep = qr(out) if qr is not None else None
            y = rec(f"layer{li}.{bi}", qa(out))
            if qr is None:
                keep = y
        feats.append(y)
    skips = feats[-2::-1]
    for i, blk in enumerate(net.decoder.blocks):
        y = F.interpolate(y, scale_factor=2, mode="nearest")
        if i < len(skips):
            y = torch.cat([y, skips[i]], dim=1)
        w1, b1 = _fold(blk.conv1[0].weight, blk.conv1[1])
        y = qa(F.relu(F.conv2d(y, qw(w1), b1, padding=1)))
        w2, b2 = _fold(blk.conv2[0].weight, blk.conv2[1])
        y = rec(f"dec{i}", qa(F.relu(F.conv2d(y, qw(w2), b2, padding=1))))
    head = net.segmentation_head[0]
    return rec("logits", F.conv2d(y, qw(head.weight), head.bias, padding=1))


@torch.no_grad()
def emulate_probs(ref_model, x, quantise=True, **kw):
    return torch.softmax(emulate_logits(ref_model, x, quantise, **kw), dim=1)

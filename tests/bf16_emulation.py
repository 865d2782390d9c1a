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


def _fold(conv_w, bn):
    s = (bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps))
    w = (conv_w.double() * s[:, None, None, None]).float()
    b = (bn.bias.double() - bn.running_mean.double() * s).float()
    return w, b


@torch.no_grad()
def emulate_logits(ref_model, x, quantise=True, taps=None):
    """`ref_model`: oracle RefUNet (eval).  x: fp32 [B,1,H,W].  Returns fp32 logits [B,C,H,W].
    `taps` (optional dict) receives intermediate activations by name."""
    q = _q if quantise else (lambda t: t)
    net = ref_model.model
    e = net.encoder

    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    w, b = _fold(e.conv1.weight, e.bn1)
    f1 = rec("stem", q(F.relu(F.conv2d(x, w, b, stride=2, padding=3))))       # stem runs in fp32, output bf16
    y = rec("pool", F.max_pool2d(f1, 3, 2, 1))
    feats = [f1]
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(e, f"layer{li}")):
            w1, b1 = _fold(blk.conv1.weight, blk.bn1)
            t = q(F.relu(F.conv2d(y, q(w1), b1, stride=blk.conv1.stride, padding=1)))
            w2, b2 = _fold(blk.conv2.weight, blk.bn2)
            acc = F.conv2d(t, q(w2), b2, padding=1)
            if blk.downsample is not None:
                wd, bd = _fold(blk.downsample[0].weight, blk.downsample[1])
                acc = acc + F.conv2d(y, q(wd), bd, stride=2)
            else:
                acc = acc + y
            y = rec(f"layer{li}.{bi}", q(F.relu(acc)))
        feats.append(y)
    skips = feats[-2::-1]
    for i, blk in enumerate(net.decoder.blocks):
        y = F.interpolate(y, scale_factor=2, mode="nearest")
        if i < len(skips):
            y = torch.cat([y, skips[i]], dim=1)
        w1, b1 = _fold(blk.conv1[0].weight, blk.conv1[1])
        y = q(F.relu(F.conv2d(y, q(w1), b1, padding=1)))
        w2, b2 = _fold(blk.conv2[0].weight, blk.conv2[1])
        y = rec(f"dec{i}", q(F.relu(F.conv2d(y, q(w2), b2, padding=1))))
    head = net.segmentation_head[0]
    return F.conv2d(y, q(head.weight), head.bias, padding=1)


@torch.no_grad()
def emulate_probs(ref_model, x, quantise=True):
    return torch.softmax(emulate_logits(ref_model, x, quantise), dim=1)

"""Parameter container with the `state_dict` layout of `smp.Unet('resnet34', in_channels=1, classes=C)`.

The reference's checkpoints (`trainer.py:46-49`) store the network under `model.<smp key>`
(`unet.py:56`); this module reproduces those keys so a reference checkpoint loads
unchanged and the engine can be fed from `state_dict()`.

Inference never runs through this module: `UNet.forward` in eval mode calls the native engine.
`forward_train` below exists only so that the reference's trainer (`unet.py:88-116`, out of scope for
this repo) can still back-propagate through the same parameters with stock PyTorch autograd.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

ENCODER_BLOCKS = {"resnet34": (3, 4, 6, 3), "resnet18": (2, 2, 2, 2)}     # torchvision BasicBlock ResNets
LAYER_BLOCKS = ENCODER_BLOCKS["resnet34"]
LAYER_CHANNELS = (64, 128, 256, 512)
DECODER_CHANNELS = (256, 128, 64, 32, 16)
SKIP_CHANNELS = (256, 128, 64, 64, 0)


def _conv(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=bias)


class _BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = _conv(cin, cout, 3, stride)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = _conv(cout, cout, 3)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class _Encoder(nn.Module):
    def __init__(self, in_channels, layer_blocks=LAYER_BLOCKS):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for i, (nb, c) in enumerate(zip(layer_blocks, LAYER_CHANNELS)):
            blocks = []
            for b in range(nb):
                blocks.append(_BasicBlock(cin, c, 2 if (i > 0 and b == 0) else 1))
                cin = c
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Sequential(_conv(cin, cout, 3), nn.BatchNorm2d(cout))
        self.conv2 = nn.Sequential(_conv(cout, cout, 3), nn.BatchNorm2d(cout))


class _Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        cins = (LAYER_CHANNELS[-1],) + DECODER_CHANNELS[:-1]
        self.blocks = nn.ModuleList(_DecoderBlock(a + s, c) for a, s, c in zip(cins, SKIP_CHANNELS, DECODER_CHANNELS))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")


class SmpUnetResnet34(nn.Module):
    """`smp.Unet(encoder_name, in_channels=1, classes=C)` for the BasicBlock ResNets (`resnet34`, the default, and
    `resnet18`: same channel widths and decoder, fewer blocks per stage)."""

    def __init__(self, in_channels=1, classes=2, encoder_name="resnet34"):
        super().__init__()
        if encoder_name not in ENCODER_BLOCKS:
            raise NotImplementedError(f"encoder {encoder_name!r}: supported are {sorted(ENCODER_BLOCKS)}")
        if in_channels != 1:
            raise NotImplementedError("the B200 engine implements num_channels=1 (uint8 grey volumes) only")
        self.encoder = _Encoder(in_channels, ENCODER_BLOCKS[encoder_name])
        self.decoder = _Decoder()
        self.segmentation_head = nn.Sequential(_conv(DECODER_CHANNELS[-1], classes, 3, bias=True))
        nn.init.xavier_uniform_(self.segmentation_head[0].weight)
        nn.init.zeros_(self.segmentation_head[0].bias)

    def forward_train(self, x):
        e = self.encoder
        feats = [F.relu(e.bn1(e.conv1(x)))]
        y = F.max_pool2d(feats[0], 3, 2, 1)
        for i in range(4):
            y = getattr(e, f"layer{i + 1}")(y)
            feats.append(y)
        skips = feats[-2::-1]
        for i, blk in enumerate(self.decoder.blocks):
            y = F.interpolate(y, scale_factor=2, mode="nearest")
            if i < len(skips):
                y = torch.cat([y, skips[i]], dim=1)
            y = F.relu(blk.conv1[1](blk.conv1[0](y)))
            y = F.relu(blk.conv2[1](blk.conv2[0](y)))
        return self.segmentation_head(y)

    def forward(self, x):
        return self.forward_train(x)

"""fp32 PyTorch restatement of the reference's network.  TEST INFRASTRUCTURE.

The reference builds its model with
`smp.Unet(encoder_name, encoder_weights, in_channels, classes)` and appends
`nn.Softmax(dim=1)` (`/root/reference/interactive_unet/unet.py:33-34,56-69`).
`segmentation-models-pytorch==0.5.0` (`pyproject.toml:19`) is a third-party
dependency that is neither vendored under `/root/reference` nor installable
here, so its published algorithm is restated for the one in-scope
configuration: `architecture='U-Net'`,
`encoder_name='resnet34'`.

Pinned parts
* Encoder: `torchvision.models.resnet34` itself (the class smp's
  `ResNetEncoder` subclasses), `fc` dropped, stem conv re-created with
  `in_channels` inputs.  Feature list `[x, relu(bn1(conv1 x)),
  layer1(maxpool .), layer2, layer3, layer4]`.
Unpinned parts ("parity unpinned": restated from smp 0.5.0's source)
* `UnetDecoder`: drop feature 0, reverse, centre = identity, five blocks of
  `nearest-interpolate to the skip's size -> cat([x, skip], 1) ->
  Conv3x3(no bias)+BN+ReLU -> Conv3x3(no bias)+BN+ReLU`, decoder channels
  (256, 128, 64, 32, 16); attention modules are identities.
* `SegmentationHead`: `Conv2d(16, classes, 3, padding=1)` with bias, no
  upsampling, no activation.
`state_dict()` keys equal smp's (`encoder.*`, `decoder.blocks.N.convM.{0,1}.*`,
`segmentation_head.0.*`); `RefUNet` prefixes them with `model.` exactly like
`unet.py:56`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

DECODER_CHANNELS = (256, 128, 64, 32, 16)
ENCODER_CHANNELS = (64, 64, 128, 256, 512)  # features 1..5 of resnet34


class _ConvBnRelu(nn.Sequential):
    # smp.base.modules.Conv2dReLU with use_norm='batchnorm': index 0 conv, 1 BN, 2 ReLU
    def __init__(self, cin, cout):
        super().__init__(
            nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(cout),
            nn.ReLU(inplace=True),
        )


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _ConvBnRelu(cin + cskip, cout)
        self.conv2 = _ConvBnRelu(cout, cout)

    def forward(self, x, size, skip=None):
        x = F.interpolate(x, size=size, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)  # upsampled tensor first, skip second
        return self.conv2(self.conv1(x))


class _Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        enc = ENCODER_CHANNELS[::-1]                       # 512, 256, 128, 64, 64
        cin = (enc[0],) + DECODER_CHANNELS[:-1]            # 512, 256, 128, 64, 32
        cskip = enc[1:] + (0,)                             # 256, 128, 64, 64, 0
        self.blocks = nn.ModuleList(
            _DecoderBlock(a, b, c) for a, b, c in zip(cin, cskip, DECODER_CHANNELS))

    def forward(self, features):
        sizes = [f.shape[2:] for f in features][::-1]      # deepest first
        feats = features[1:][::-1]
        x, skips = feats[0], feats[1:]
        for i, blk in enumerate(self.blocks):
            skip = skips[i] if i < len(skips) else None
            x = blk(x, tuple(sizes[i + 1]), skip)
        return x


class RefSmpUnetResnet34(nn.Module):
    """`smp.Unet(encoder_name, encoder_weights=None, in_channels, classes)` for `resnet34` (default) or `resnet18`
    (same feature widths, so the same decoder; smp's `resnet18` encoder is torchvision's class as well)."""

    def __init__(self, in_channels=1, classes=2, encoder_name="resnet34"):
        super().__init__()
        if encoder_name not in ("resnet34", "resnet18"):
            raise NotImplementedError(encoder_name)
        enc = getattr(torchvision.models, encoder_name)(weights=None)
        del enc.fc                                          # smp: `del self.fc`
        enc.avgpool = nn.Identity()                         # never called
        if in_channels != 3:
            enc.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
            nn.init.kaiming_normal_(enc.conv1.weight, mode="fan_out", nonlinearity="relu")
        self.encoder = enc
        self.decoder = _Decoder()
        self.segmentation_head = nn.Sequential(
            nn.Conv2d(DECODER_CHANNELS[-1], classes, kernel_size=3, padding=1),
            nn.Identity(), nn.Identity())
        # smp initialises decoder convs with kaiming_uniform(fan_in, relu) and the head with xavier_uniform
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
        nn.init.xavier_uniform_(self.segmentation_head[0].weight)
        nn.init.constant_(self.segmentation_head[0].bias, 0)

    def encode(self, x):
        e = self.encoder
        f0 = x
        f1 = e.relu(e.bn1(e.conv1(x)))
        f2 = e.layer1(e.maxpool(f1))
        f3 = e.layer2(f2)
        f4 = e.layer3(f3)
        f5 = e.layer4(f4)
        return [f0, f1, f2, f3, f4, f5]

    def forward(self, x):
        if x.shape[2] % 32 or x.shape[3] % 32:
            raise RuntimeError(f"Wrong input shape height={x.shape[2]}, width={x.shape[3]}. Expected image "
                               "height and width divisible by 32.")
        return self.segmentation_head(self.decoder(self.encode(x)))

    def state_dict(self, *a, **k):
        sd = super().state_dict(*a, **k)
        return type(sd)((n, v) for n, v in sd.items() if not n.startswith("encoder.avgpool"))


class RefUNet(nn.Module):
    """The reference's `UNet` (`unet.py:10-69`) minus Lightning: probabilities out."""

    def __init__(self, num_channels=1, num_classes=2, encoder_name="resnet34"):
        super().__init__()
        self.model = RefSmpUnetResnet34(num_channels, num_classes, encoder_name)
        self.softmax = nn.Softmax(dim=1)                    # unet.py:63

    def forward(self, x):
        return self.softmax(self.model(x))                  # unet.py:67

    @property
    def device(self):
        return next(self.parameters()).device


def conv_macs_per_slice(size, classes=2, encoder_name="resnet34"):
    """Dense multiply-accumulates of every conv for one `size` x `size` slice."""
    net = RefSmpUnetResnet34(1, classes, encoder_name).eval()
    total = [0]

    def hook(m, inp, out):
        k = m.kernel_size[0] * m.kernel_size[1]
        total[0] += out.shape[2] * out.shape[3] * k * m.in_channels * m.out_channels

    hs = [m.register_forward_hook(hook) for m in net.modules() if isinstance(m, nn.Conv2d)]
    with torch.inference_mode():
        net(torch.zeros(1, 1, size, size))
    for h in hs:
        h.remove()
    return total[0]

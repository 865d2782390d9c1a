"""Writes smp_unet_resnet34_keys.json: the (key, shape) list of `smp.Unet('resnet34', in_channels=1, classes=2)
.state_dict()` as segmentation-models-pytorch 0.5.0 lays it out (the reference pins that version, pyproject.toml:19).

smp itself is not installable here, so the list is assembled from two independent sources instead of from this
repository's own module: the ENCODER half is read off `torchvision.models.resnet34()` (smp's ResNetEncoder is that class
minus `fc`, with conv1 re-shaped to one input channel), the DECODER / HEAD half is written out from smp 0.5.0's published
`UnetDecoder` (`DecoderBlock.conv1 / conv2` = `Conv2dReLU` = Sequential(conv, BatchNorm2d, ReLU); `center` and the
attention modules are Identity for this configuration) and `SegmentationHead` (Sequential(conv 3x3 with bias,
Identity, Identity)).  Run:  python tests/golden/make_smp_keys.py
"""
import json
import os

import torchvision

BN = (("weight", None), ("bias", None), ("running_mean", None), ("running_var", None), ("num_batches_tracked", ()))


def main(classes=2):
    keys = []
    for k, v in torchvision.models.resnet34().state_dict().items():
        if k.startswith("fc."):
            continue
        shape = list(v.shape)
        if k == "conv1.weight":
            shape[1] = 1                                     # in_channels=1 (smp patches the first conv)
        keys.append(["encoder." + k, shape])
    encoder_out = [64, 64, 128, 256, 512]                    # features at 1/2 .. 1/32 (the 1/1 identity feature has 1)
    dec_out = [256, 128, 64, 32, 16]                         # decoder_channels default
    skips = encoder_out[::-1][1:] + [0]                      # 256, 128, 64, 64, then no skip for the last block
    cin = encoder_out[-1]
    for i, (s, co) in enumerate(zip(skips, dec_out)):
        for name, ci in (("conv1", cin + s), ("conv2", co)):
            keys.append([f"decoder.blocks.{i}.{name}.0.weight", [co, ci, 3, 3]])
            for b, shp in BN:
                keys.append([f"decoder.blocks.{i}.{name}.1.{b}", [co] if shp is None else list(shp)])
        cin = co
    keys.append(["segmentation_head.0.weight", [classes, dec_out[-1], 3, 3]])
    keys.append(["segmentation_head.0.bias", [classes]])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "smp_unet_resnet34_keys.json")
    json.dump({"model": "smp.Unet('resnet34', in_channels=1, classes=2), segmentation-models-pytorch 0.5.0",
               "keys": keys}, open(path, "w"), indent=0)
    print(len(keys), "keys ->", path)


if __name__ == "__main__":
    main()

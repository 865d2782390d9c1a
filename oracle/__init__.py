"""CPU oracle for the volume-prediction path of laprade117/interactive-unet.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package
(`interactive-unet_b200/`, imported as `interactive_unet_b200`) may import
this directory.  The only legal callers are `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`, and they use it as the checker or as the timed CPU baseline,
never as the product path.

Contents
--------
reference_loader   imports `/root/reference/interactive_unet/predict.py`
                   VERBATIM (with `sys.modules` stubs for the five packages
                   that are absent from this image).  Only works where
                   `/root/reference` exists (the build container); it is what
                   the restatements below are pinned against and what
                   `make_golden.py` records fixtures from.
smp_unet_resnet34  plain-PyTorch fp32 restatement of
                   `segmentation-models-pytorch==0.5.0`'s `Unet('resnet34')`
                   (third-party, not vendored in the reference) wrapped the
                   way `unet.py:56-69` wraps it.
predict_port       numpy restatement of `predict.py`'s slice / accumulate /
                   blend / quantise arithmetic, usable on the GPU box where
                   `/root/reference` does not exist.
synth              seeded weights and volumes.
make_golden        writes `tests/golden/*.npz` from the verbatim reference.

Parity status
-------------
The reference ships no tests and no golden vectors, so:
* everything in `predict_port` is PINNED against the verbatim reference
  functions executed in this container (fixtures under `tests/golden/`,
  generator `oracle/make_golden.py`);
* the encoder half of the network is pinned against
  `torchvision.models.resnet34`;
* the decoder / head half restates smp 0.5.0 from its published source and
  is "parity unpinned" (the package is not installable here).
"""

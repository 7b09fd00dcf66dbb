"""Backbones for drop-in runs WITHOUT network access (random-init weights only).

The backbone forward is not part of the accelerated path (it stays in PyTorch and is timed apart);
these exist so that the reference's call sequence `backbones.load(name)` ->
`AnomalyClusteringCore.load(backbone, layers_to_extract_from, ...)` can run end to end here.
Reference: models/patchcore/backbones.py:53-79 (name -> constructor table; it downloads DINO weights,
which is impossible offline) and models/patchcore/vision_transformer.py:134-254 (DINO ViT)."""
from __future__ import annotations

import torch

from .vit import VisionTransformer, vit_base, vit_small  # noqa: F401


class RandomInitBackboneWarning(UserWarning):
    pass


def load(name: str, seed: int = 2023, allow_random_init: bool = False) -> torch.nn.Module:
    """name: 'wideresnet50' | 'dino_vitbase8' | 'dino_vitsmall8' | 'dino_vitbase16'.

    The networks are RANDOM-INIT (the reference's `patchcore.backbones.load` downloads pretrained DINO / ImageNet
    weights, backbones.py:56-79, which is impossible offline): alpha / X / NMI computed from them are meaningless for
    real data.  The caller must say so explicitly (`allow_random_init=True`); a warning is emitted either way and the
    returned module carries `random_init = True` so that drivers can tag their output."""
    import warnings

    if not allow_random_init:
        raise ValueError("backbones.load(%r) builds a RANDOM-INIT network (no pretrained weights offline). Pass a pretrained "
                         "torch module as `backbone=` instead, or allow_random_init=True for shape / smoke runs." % name)
    warnings.warn("backbone %r is RANDOM-INIT: results are for shape / performance checks only, not for real data" % name,
                  RandomInitBackboneWarning, stacklevel=2)
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        if name == "wideresnet50":
            import torchvision.models as models

            net = models.wide_resnet50_2(weights=None)
        elif name == "dino_vitbase8":
            net = vit_base(patch_size=8)
        elif name == "dino_vitbase16":
            net = vit_base(patch_size=16)
        elif name == "dino_vitsmall8":
            net = vit_small(patch_size=8)
        else:
            raise KeyError("unknown backbone %r (offline: wideresnet50, dino_vitbase8, dino_vitbase16, dino_vitsmall8)" % name)
    finally:
        torch.random.set_rng_state(gen_state)
    net.name = name
    net.random_init = True
    return net.eval()

"""Synthetic inputs for parity tests and benchmarks (no MVTec, no pretrained weights here).

Feature-level generator: per layer a class template (fixed random low-rank
field) + N(0,1) noise + planted 'defect' blocks, so that Ward clustering has real structure and
the patch distances have a realistic spread.  Seeds follow the reference's same_seeds(2023)
(examples/main.py:62-69)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def planted_features(
    n_img: int,
    layers: Sequence[Tuple[int, int, int, bool]],
    n_classes: int = 4,
    seed: int = 2023,
    device="cpu",
    defect_gain: float = 3.0,
    noise: float = 1.0,
    channel_bias: float = 0.0,
):
    """layers: [(C, H, W, tokens)] -> (list of per-layer feature tensors, labels [n_img]).

    tokens=True yields ViT block outputs [N, 1+H*W, C] (CLS first), else CNN maps [N,C,H,W].
    Class 0 is 'good'; class c>0 plants a defect block whose channel signature depends on c.
    channel_bias > 0 adds a fixed per-channel offset N(0, channel_bias^2) to every image and position (real
    ViT features carry strong per-channel means): it survives the pooling, so the patch norms grow (45-50 at
    1.45 for the config-2 shape) while the patch DISTANCES shrink -- the hard case for the |x|^2+|y|^2-2xy form."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    labels = torch.arange(n_img) % n_classes
    # defect geometry is shared across layers (relative coordinates)
    rel = torch.rand(n_img, 2, generator=gen)
    size = torch.randint(3, 7, (n_img,), generator=gen)
    feats: List[torch.Tensor] = []
    for (C, H, W, tokens) in layers:
        rank = 8
        u = torch.randn(C, rank, generator=gen)
        v = torch.randn(rank, H * W, generator=gen)
        template = (u @ v).reshape(1, C, H, W) / rank ** 0.5
        x = template + noise * torch.randn(n_img, C, H, W, generator=gen)
        sig = torch.randn(n_classes, C, generator=gen)
        if channel_bias:
            x = x + channel_bias * torch.randn(C, generator=gen).reshape(1, C, 1, 1)
        for i in range(n_img):
            c = int(labels[i])
            if c == 0:
                continue
            sh = max(1, int(size[i]) * H // 28)
            sw = max(1, int(size[i]) * W // 28)
            y0 = int(rel[i, 0] * (H - sh))
            x0 = int(rel[i, 1] * (W - sw))
            x[i, :, y0 : y0 + sh, x0 : x0 + sw] += defect_gain * sig[c].reshape(C, 1, 1)
        if tokens:
            t = x.permute(0, 2, 3, 1).reshape(n_img, H * W, C)
            cls = torch.randn(n_img, 1, C, generator=gen)
            x = torch.cat([cls, t], dim=1)
        feats.append(x.contiguous().to(device))
    return feats, labels


CONFIGS = {
    # BASELINE.json configs[0]: WideResNet50 layer2+layer3, 20 images, 1024 -> 1024
    "config1": dict(layers=[(512, 28, 28, False), (1024, 14, 14, False)], n_img=20, Dp=1024, D=1024, classes=4),
    # configs[1]: DINO ViT-B/8 blocks.10+blocks.11, ~100 images, 2048 -> 4096
    "config2": dict(layers=[(768, 28, 28, True), (768, 28, 28, True)], n_img=100, Dp=2048, D=4096, classes=4),
    # configs[4]: ViT-S/8 at 448x448 (3136 patches/img)
    "config5": dict(layers=[(384, 56, 56, True), (384, 56, 56, True)], n_img=64, Dp=2048, D=4096, classes=4),
}


def planted_features_device(
    img_ids: Sequence[int],
    layers: Sequence[Tuple[int, int, int, bool]],
    n_classes: int = 4,
    seed: int = 2023,
    device="cuda",
    defect_gain: float = 3.0,
    channel_bias: float = 0.0,
):
    """Same construction as planted_features, generated directly on the device per GLOBAL image id
    (so any rank can produce exactly its slice of a fixed synthetic data set).  Used by bench.py."""
    dev = torch.device(device)
    feats: List[torch.Tensor] = []
    ids = list(img_ids)
    for li, (C, H, W, tokens) in enumerate(layers):
        g = torch.Generator(device=dev).manual_seed(seed * 1000 + li)
        rank = 8
        u = torch.randn(C, rank, generator=g, device=dev)
        v = torch.randn(rank, H * W, generator=g, device=dev)
        template = (u @ v).reshape(C, H, W) / rank ** 0.5
        sig = torch.randn(n_classes, C, generator=g, device=dev)
        if channel_bias:
            template = template + channel_bias * torch.randn(C, generator=g, device=dev).reshape(C, 1, 1)
        T = (1 + H * W) if tokens else 0
        out = torch.empty((len(ids), T, C) if tokens else (len(ids), C, H, W), dtype=torch.float32, device=dev)
        for n, i in enumerate(ids):
            gi = torch.Generator(device=dev).manual_seed(seed * 100003 + i * 17 + li)
            x = template + torch.randn(C, H, W, generator=gi, device=dev)
            c = i % n_classes
            if c != 0:
                gc = torch.Generator().manual_seed(seed * 7919 + i)
                rel = torch.rand(2, generator=gc)
                size = int(torch.randint(3, 7, (1,), generator=gc))
                sh = max(1, size * H // 28)
                sw = max(1, size * W // 28)
                y0 = int(rel[0] * (H - sh))
                x0 = int(rel[1] * (W - sw))
                x[:, y0 : y0 + sh, x0 : x0 + sw] += defect_gain * sig[c].reshape(C, 1, 1)
            if tokens:
                out[n, 0] = torch.randn(C, generator=gi, device=dev)
                out[n, 1:] = x.permute(1, 2, 0).reshape(H * W, C)
            else:
                out[n] = x
        feats.append(out)
    labels = torch.tensor([i % n_classes for i in ids])
    return feats, labels

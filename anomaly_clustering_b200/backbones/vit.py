"""A plain pre-LN Vision Transformer with the module naming the reference's hooks rely on
(`blocks.<i>` outputs are [B, 1+P, C] token tensors with the CLS token first; the final `norm` is
not applied to hooked block outputs).  Written from the standard ViT definition; interpolates the
position embedding for other input sizes (needed for BASELINE config 5: ViT-S/8 at 448x448)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, T, C = x.shape
        qkv = self.qkv(x).reshape(B, T, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        return self.proj(out.transpose(1, 2).reshape(B, T, C))


class Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = nn.Sequential(nn.Linear(dim, int(dim * mlp_ratio)), nn.GELU(), nn.Linear(int(dim * mlp_ratio), dim))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=8, embed_dim=768, depth=12, heads=12):
        super().__init__()
        self.patch_size = patch_size
        self.patch_embed = nn.Conv2d(3, embed_dim, kernel_size=patch_size, stride=patch_size)
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)

    def _pos(self, h, w):
        n = self.pos_embed.shape[1] - 1
        s = int(math.sqrt(n))
        if h * w == n and h == w:
            return self.pos_embed
        grid = self.pos_embed[:, 1:].reshape(1, s, s, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=(h, w), mode="bicubic", align_corners=False)
        return torch.cat([self.pos_embed[:, :1], grid.permute(0, 2, 3, 1).reshape(1, h * w, -1)], dim=1)

    def forward(self, x):
        B = x.shape[0]
        x = self.patch_embed(x)
        h, w = x.shape[-2:]
        x = x.flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(B, -1, -1), x], dim=1) + self._pos(h, w)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)[:, 0]


def vit_small(patch_size=8, **kw):
    return VisionTransformer(patch_size=patch_size, embed_dim=384, depth=12, heads=6, **kw)


def vit_base(patch_size=8, **kw):
    return VisionTransformer(patch_size=patch_size, embed_dim=768, depth=12, heads=12, **kw)

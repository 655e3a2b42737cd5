"""DINOv2 ViT-L/14 parameter container + forward (row a1 of SURVEY.md §8).

The reference obtains this module from `torch.hub.load("facebookresearch/dinov2", "dinov2_vitl14")`
(configs/model/ae_net/dinov2_l.yaml:4-7) and only ever calls `forward_features(x)["x_prenorm"]`
(ae_net.py:46,65).  This class carries the SAME state-dict keys as upstream's DinoVisionTransformer
(`cls_token, pos_embed, mask_token, patch_embed.proj.*, blocks.{i}.{norm1,attn.qkv,attn.proj,ls1.gamma,norm2,
mlp.fc1,mlp.fc2,ls2.gamma}.*, norm.*`), so `gigaPose_v1.ckpt` loads strictly into `ae_net.dinov2_model.*`.
`AENet` accepts either this class or a hub module and runs the forward through `gigapose_b200.vit_engine`.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _LayerScale(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, heads, ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.ls1 = _LayerScale(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, dim * ratio)
        self.ls2 = _LayerScale(dim)


class DinoVisionTransformer(nn.Module):
    def __init__(self, embed_dim=1024, depth=24, num_heads=16, patch_size=14, mlp_ratio=4, train_grid=37,
                 init_seed: Optional[int] = None):
        super().__init__()
        self.embed_dim, self.depth, self.num_heads, self.patch_size = embed_dim, depth, num_heads, patch_size
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + train_grid * train_grid, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self.patch_embed = _PatchEmbed(embed_dim, patch_size)
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        if init_seed is not None:
            self.seeded_init(init_seed)

    @torch.no_grad()
    def seeded_init(self, seed: int) -> None:
        """Deterministic non-degenerate weights for synthetic runs (no checkpoint is reachable offline)."""
        g = torch.Generator().manual_seed(seed)
        for name, p in self.named_parameters():
            if name.endswith("gamma"):
                p.fill_(1.0)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif p.dim() >= 2 and "pos_embed" not in name and "token" not in name:
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(p[0].numel()))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))

    def interpolated_pos_embed(self, gh: int, gw: int) -> torch.Tensor:
        """Weight-only: the [1, 1+gh*gw, dim] table upstream rebuilds at every forward (bicubic, offset 0.1)."""
        pe = self.pos_embed.float()
        n = pe.shape[1] - 1
        m = int(math.isqrt(n))
        if gh * gw == n and gh == gw:
            return pe
        patch = pe[:, 1:].reshape(1, m, m, self.embed_dim).permute(0, 3, 1, 2)
        patch = F.interpolate(patch, scale_factor=((gh + 0.1) / m, (gw + 0.1) / m), mode="bicubic")
        patch = patch.permute(0, 2, 3, 1).reshape(1, gh * gw, self.embed_dim)
        return torch.cat([pe[:, :1], patch], dim=1)

    def forward_features(self, x: torch.Tensor):
        from .vit_engine import vit_forward_features
        return {"x_prenorm": vit_forward_features(self, x)}

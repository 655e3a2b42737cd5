"""Executes the ViT-L/14 forward for `AENet` (row a1).

INTERIM (round 1): the transformer blocks still run through torch's library kernels (cuBLAS GEMMs + SDPA) in
fp32 -- this stage is explicitly listed as `library` in bench.py's config and in DESIGN.md; the hand-written
tcgen05 GEMM / attention kernels replace it next (DESIGN.md "what comes next").  Everything downstream of the
patch tokens (normalisation, bank layout, similarity search, IST MLP, RANSAC, pose) is native.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BACKEND = "torch-library"


def _weights_of(model):
    """Works for gigapose_b200.vit.DinoVisionTransformer and for upstream hub modules (same attribute names)."""
    return model


@torch.no_grad()
def vit_forward_features(model, x: torch.Tensor) -> torch.Tensor:
    """x [b,3,H,W] -> x_prenorm [b, 1+gh*gw, dim] (tokens after the last block, before the final norm)."""
    m = _weights_of(model)
    B, _, H, W = x.shape
    ps = m.patch_size if isinstance(m.patch_size, int) else m.patch_size[0]
    gh, gw = H // ps, W // ps
    if hasattr(m, "interpolated_pos_embed"):
        pos = m.interpolated_pos_embed(gh, gw)
    else:                                     # upstream module: its own (identical) routine
        pos = m.interpolate_pos_encoding(torch.zeros(1, 1 + gh * gw, m.pos_embed.shape[-1], device=x.device), W, H)
    tok = F.conv2d(x, m.patch_embed.proj.weight, m.patch_embed.proj.bias, stride=ps).flatten(2).transpose(1, 2)
    tok = torch.cat([m.cls_token.expand(B, -1, -1), tok], dim=1) + pos.to(tok.dtype)
    for blk in m.blocks:
        h = F.layer_norm(tok, (tok.shape[-1],), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        nh = blk.attn.num_heads
        qkv = F.linear(h, blk.attn.qkv.weight, blk.attn.qkv.bias).reshape(B, -1, 3, nh, h.shape[-1] // nh)
        q, k, v = qkv.permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, -1, h.shape[-1])
        tok = tok + blk.ls1.gamma * F.linear(a, blk.attn.proj.weight, blk.attn.proj.bias)
        h = F.layer_norm(tok, (tok.shape[-1],), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
        h = F.linear(F.gelu(F.linear(h, blk.mlp.fc1.weight, blk.mlp.fc1.bias)), blk.mlp.fc2.weight, blk.mlp.fc2.bias)
        tok = tok + blk.ls2.gamma * h
    return tok

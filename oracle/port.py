"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the GigaPose inference hot path.

This module is the parity oracle that travels to the GPU box (the reference tree does not).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product path (``gigapose_b200``/``src``) never does.

Pinning: the reference ships no tests or golden vectors, so the port is pinned against
outputs of the *unmodified reference modules run in the build container* (``oracle/ref_import.py``):
``oracle/make_golden.py`` writes ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` requires this
port to reproduce them (bit-exact on integer tensors, <=1e-6 on floats).  When the reference tree is
present the same test also compares live.

Every function cites the reference file:line it restates.  Plain fp32 torch on CPU, same operation order
as the reference wherever the order changes bits (einsum contraction, mask multiplication order,
threshold, first-max tie-breaking).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

G = 16  # patches per side


# ------------------------------------------------------------------------------------------------------------
# a4  LocalSimilarity.test        (src/models/matching.py:188-316; helpers :29-113)
# ------------------------------------------------------------------------------------------------------------
def similarity_search(src_feats, tar_feat, src_masks, tar_mask, k=5, sim_threshold=0.5, patch_threshold=3,
                      chunk=32, return_intermediates=False) -> Dict[str, torch.Tensor]:
    """src_feats [B,N,C,16,16], tar_feat [B,C,16,16], src_masks [B,N,224,224], tar_mask [B,224,224]."""
    outs = {n: [] for n in ("id_src", "score_src", "score_pts", "tar_pts", "src_pts")}
    inter = {n: [] for n in ("sim_avg", "idx_tar2src", "idx_src2tar", "score_tar2src", "score_src2tar", "mask_all")}
    Btot = src_feats.shape[0]
    for b0 in range(0, Btot, chunk):                                          # matching.py:212 (BatchedData chunks)
        sf, tf = src_feats[b0:b0 + chunk], tar_feat[b0:b0 + chunk]
        sm, tm = src_masks[b0:b0 + chunk], tar_mask[b0:b0 + chunk]
        B, N = sm.shape[:2]
        # matching.py:222-230 -- nearest 224->16 sampling of the masks, second L2 normalisation of the features
        tm = F.interpolate(tm.unsqueeze(1), size=(G, G)).reshape(B, G * G)
        tf = F.normalize(tf, dim=1).reshape(B, tf.shape[1], G * G)
        sm = F.interpolate(sm, size=(G, G)).reshape(B, N, G * G)
        sf = F.normalize(sf, dim=2).reshape(B, N, sf.shape[2], G * G)
        # matching.py:233-236
        sim = torch.einsum("b c t, b n c s -> b n t s", tf, sf)
        sim *= sm[:, :, None, :]
        sim *= tm[:, None, :, None]
        sim[sim < sim_threshold] = 0
        # matching.py:240-241 (tar2src direction)
        score_t2s, idx_t2s = torch.max(sim, dim=3)
        score_s2t, idx_s2t = torch.max(sim, dim=2)
        mask_sim = score_t2s >= sim_threshold                                  # :247
        # matching.py:80-113 cycle consistency
        back = torch.gather(idx_s2t, 2, idx_t2s)
        bx, by = (back % G).float(), (back // G).float()
        t_ids = torch.arange(G * G)
        gx, gy = (t_ids % G).float(), (t_ids // G).float()
        dist = torch.norm(torch.stack([bx - gx, by - gy], dim=-1), dim=3)
        mask_cycle = torch.logical_and(dist <= patch_threshold,
                                       torch.gather(score_s2t, 2, idx_t2s) >= sim_threshold)
        # matching.py:259-268 (incl. the s/t index mix-up of `idx_src2tar != 0`)
        mask_non_zero = tm[:, None, :].expand(B, N, G * G) * torch.gather(sm, 2, idx_t2s) \
            * (idx_s2t != 0) * (idx_t2s != 0)
        mask_all = mask_sim * mask_cycle * mask_non_zero                        # :271
        # matching.py:274-279
        any_valid = mask_all.sum(dim=2) > 0
        sim_avg = torch.zeros(B, N)
        sim_avg[any_valid] = torch.sum(score_t2s * mask_all, dim=2)[any_valid] / (G * G)
        score_src, id_src = torch.topk(sim_avg, k, dim=1)
        # matching.py:282-300 + format_prediction :29-61
        bi = torch.arange(B)[:, None].expand(B, k)
        win_mask = mask_all[bi, id_src]
        win_idx = idx_t2s[bi, id_src]
        valid = win_mask != 0
        minus1 = torch.full((B, k, G * G), -1, dtype=torch.long)
        tx = torch.where(valid, (t_ids % G).expand(B, k, G * G), minus1)
        ty = torch.where(valid, (t_ids // G).expand(B, k, G * G), minus1)
        sx = torch.where(valid, win_idx % G, minus1)
        sy = torch.where(valid, win_idx // G, minus1)
        outs["id_src"].append(id_src)
        outs["score_src"].append(score_src)
        outs["score_pts"].append(score_t2s[bi, id_src])
        outs["tar_pts"].append(torch.stack([tx, ty], dim=-1))
        outs["src_pts"].append(torch.stack([sx, sy], dim=-1))
        if return_intermediates:
            for n, v in (("sim_avg", sim_avg), ("idx_tar2src", idx_t2s), ("idx_src2tar", idx_s2t),
                         ("score_tar2src", score_t2s), ("score_src2tar", score_s2t), ("mask_all", mask_all)):
                inter[n].append(v)
    res = {n: torch.cat(v, 0) for n, v in outs.items()}
    if return_intermediates:
        res.update({n: torch.cat(v, 0) for n, v in inter.items()})
    return res


# ------------------------------------------------------------------------------------------------------------
# a5  ISTNet.inference + gather + Regressor   (ist_net.py:97-162, utils/batch.py:46-73)
# ------------------------------------------------------------------------------------------------------------
class RegressorPort(nn.Module):
    """Same parameter names as the reference Regressor (ist_net.py:123-162)."""

    def __init__(self, descriptor_size=256, hidden_dim=256, use_tanh_act=True, seed: Optional[int] = 9):
        super().__init__()
        d, h = descriptor_size, hidden_dim
        self.scale_predictor = nn.Sequential(nn.Linear(2 * d, 2 * h), nn.ReLU(), nn.Linear(2 * h, h), nn.ReLU(),
                                             nn.Linear(h, 1))
        self.inplane_predictor = nn.Sequential(nn.Linear(2 * d, 2 * h), nn.ReLU(), nn.Linear(2 * h, h), nn.ReLU(),
                                               nn.Linear(h, 2), nn.Tanh() if use_tanh_act else nn.Identity())
        if seed is not None:
            g = torch.Generator().manual_seed(seed)
            with torch.no_grad():
                for name, p in self.named_parameters():
                    if p.dim() == 2:
                        p.copy_(torch.randn(p.shape, generator=g) * math.sqrt(2.0 / p.shape[1]))
                    else:
                        p.copy_(0.05 * torch.randn(p.shape, generator=g))


def _pick(features, pts):
    """utils/batch.py:46-73 without the final boolean compaction: features [B,C,16,16], pts [B,N,2] (x,y)."""
    B, C, H, W = features.shape
    valid = (pts[..., 0] != -1) & (pts[..., 1] != -1)
    p = pts.clone()
    p[p == -1] = H - 1
    idx = p[..., 1] * W + p[..., 0]
    flat = features.reshape(B, C, H * W).permute(0, 2, 1)
    return torch.gather(flat, 1, idx[..., None].expand(-1, -1, C)), valid


@torch.no_grad()
def ist_mlp(regressor, src_feat, tar_feat, src_pts, tar_pts):
    """ist_net.py:97-120: rows = cat(tar feature at tar_pt, src feature at src_pt); -1000 fill elsewhere."""
    fs, vs = _pick(src_feat, src_pts)
    ft, vt = _pick(tar_feat, tar_pts)
    assert int(vs.sum()) == int(vt.sum())
    rows = torch.cat([ft[vt], fs[vs]], dim=1)
    B, N = src_pts.shape[:2]
    scales = torch.full((B, N), -1000.0)
    cs = torch.full((B, N, 2), -1000.0)
    if rows.shape[0]:
        scales[vs] = regressor.scale_predictor(rows).squeeze(1)
        cs[vs] = regressor.inplane_predictor(rows)
    return scales, cs


# ------------------------------------------------------------------------------------------------------------
# a7  exhaustive one-point RANSAC   (ransac.py:19-172, poses.py:124-163, lib3d/torch.py:7-89)
# ------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def ransac(src_pts, tar_pts, rel_scale, rel_inplane, pixel_threshold=14.0, patch_size=14):
    """src_pts/tar_pts [B,K,256,2] i64 (-1 invalid), rel_scale [B,K,256], rel_inplane [B,K,256,2] (cos,sin).
    Returns M [B,K,3,3] f32, failed [B,K] bool, inlier src/tar pts [B,K,256,2] i64, inlier scores [B,K,256] i64."""
    B, K, N = src_pts.shape[:3]
    Ms = torch.eye(3).repeat(B, K, 1, 1)
    failed = torch.zeros(B, K, dtype=torch.bool)
    in_src = torch.full((B, K, N, 2), -1, dtype=torch.long)
    in_tar = torch.full((B, K, N, 2), -1, dtype=torch.long)
    in_sc = torch.zeros(B, K, N, dtype=torch.long)
    for b in range(B):
        for kk in range(K):
            keep = src_pts[b, kk, :, 0] != -1                                     # ransac.py:141
            n = int(keep.sum())
            if n < 1:
                continue
            s_i, t_i = src_pts[b, kk][keep], tar_pts[b, kk][keep]
            s = (s_i * patch_size).float()                                        # :57-58 pixel units, no half-patch offset
            t = (t_i * patch_size).float()
            sc, cs = rel_scale[b, kk][keep], rel_inplane[b, kk][keep]
            c, sn = cs[:, 0], cs[:, 1]
            # affine_torch (lib3d/torch.py:21-29): 2x2 block = rotation * scale, elementwise
            m00, m01, m10, m11 = c * sc, (-sn) * sc, sn * sc, c * sc
            M = torch.eye(3).repeat(n, 1, 1)
            M[:, 0, 0], M[:, 0, 1], M[:, 1, 0], M[:, 1, 1] = m00, m01, m10, m11
            # apply_affine on the proposing point (translation still zero), ransac.py:91-93
            h = torch.cat([s, torch.ones(n, 1)], dim=1)
            a = torch.einsum("bhc,bc->bh", M, h)
            a = a[:, :2] / a[:, 2:]
            M[:, :2, 2] = t - a
            # score every candidate on all other correspondences, ransac.py:96-101
            allp = torch.einsum("bhc,nc->bnh", M, h)
            allp = allp[:, :, :2] / allp[:, :, 2:]
            err = torch.norm(t[None] - allp, dim=2)
            inl = err <= pixel_threshold
            inl[torch.arange(n), torch.arange(n)] = False                         # validation set excludes the proposer (:29-33)
            score = inl.sum(dim=1)
            best_score, best = torch.max(score, dim=0)
            failed[b, kk] = best_score == 0
            Ms[b, kk] = M[best]
            idx = torch.where(inl[best])[0]
            in_src[b, kk, : len(idx)] = s_i[idx]
            in_tar[b, kk, : len(idx)] = t_i[idx]
            in_sc[b, kk, : len(idx)] = 1
    return Ms, failed, in_src, in_tar, in_sc


# ------------------------------------------------------------------------------------------------------------
# a9  pose lifting   (poses.py:26-122, lib3d/torch.py:47-65,150-162)
# ------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def pose_recovery(tar_label, tar_K, tar_M, id_src, pred_M, template_K, template_Ms, template_poses):
    B, K = id_src.shape
    lab = tar_label - 1
    tK = template_K[lab][:, None].expand(B, K, 3, 3)
    bi = torch.arange(B)[:, None].expand(B, K)
    tM = template_Ms[lab][bi, id_src]
    poses = template_poses[lab][bi, id_src].clone()
    # normalize_affine_transform (lib3d/torch.py:150-162)
    scale = torch.norm(pred_M[:, :, :2, 0], dim=2)
    Rin = torch.zeros(B, K, 3, 3)
    Rin[:, :, 2, 2] = 1
    Rin[:, :, :2, :2] = pred_M[:, :, :2, :2] / scale[:, :, None, None]
    poses[:, :, :3, :3] = torch.matmul(Rin, poses[:, :, :3, :3])                  # poses.py:69-71
    temp_z = poses[:, :, 2, 3].clone()
    c2d = torch.matmul(tK, poses[:, :, :3, 3].unsqueeze(-1))
    c2d = c2d / c2d[:, :, 2].unsqueeze(2)
    # inverse_affine of the (scale+translation) query crop matrix (lib3d/torch.py:47-65)
    qs = tar_M[:, 0, 0]
    Minv = torch.eye(3).repeat(B, 1, 1)
    Minv[:, 0, 0] = 1 / qs
    Minv[:, 1, 1] = 1 / qs
    Minv[:, :2, 2] = -tar_M[:, :2, 2] / qs.unsqueeze(1)
    aff = torch.matmul(torch.matmul(Minv[:, None].expand(B, K, 3, 3), pred_M), tM)   # poses.py:83
    qc = torch.matmul(aff, c2d)
    qK = tar_K[:, None].expand(B, K, 3, 3)
    qKinv = torch.inverse(qK)
    s2d = torch.norm(aff[:, :, :2, 0], dim=2)
    qz = (temp_z / s2d) * (qK[:, :, 0, 0] / tK[:, :, 0, 0])
    tr = torch.matmul(qKinv, qc).squeeze(-1)
    tr = tr / tr[:, :, 2].unsqueeze(-1)
    poses[:, :, :3, 3] = tr * qz.unsqueeze(-1)
    return poses


# ------------------------------------------------------------------------------------------------------------
# a1  DINOv2 ViT-L/14 forward_features()["x_prenorm"]   (ae_net.py:46,55-69; upstream facebookresearch/dinov2,
#     un-vendored & un-pinned -- architecture restated from SURVEY.md Appendix B; parity for the ViT is
#     therefore pinned only against this restatement + the HF `transformers` Dinov2 port cross-check)
# ------------------------------------------------------------------------------------------------------------
class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * (C // self.num_heads) ** -0.5, qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((a @ v).transpose(1, 2).reshape(B, N, C))


class _LS(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, heads, ratio=4):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, heads)
        self.ls1 = _LS(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, dim * ratio)
        self.ls2 = _LS(dim)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class DinoV2Port(nn.Module):
    """State-dict keys follow upstream DinoVisionTransformer (cls_token, pos_embed, mask_token,
    patch_embed.proj.*, blocks.{i}.{norm1,attn.qkv,attn.proj,ls1.gamma,norm2,mlp.fc1,mlp.fc2,ls2.gamma}.*, norm.*)."""

    def __init__(self, dim=1024, depth=24, heads=16, patch=14, train_grid=37, seed: Optional[int] = 7):
        super().__init__()
        self.patch_size, self.dim = patch, dim
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + train_grid * train_grid, dim))
        self.mask_token = nn.Parameter(torch.zeros(1, dim))
        self.patch_embed = _PatchEmbed(dim, patch)
        self.blocks = nn.ModuleList([_Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        if seed is not None:
            self.seeded_init(seed)

    @torch.no_grad()
    def seeded_init(self, seed):
        g = torch.Generator().manual_seed(seed)
        for name, p in self.named_parameters():
            if name.endswith("gamma"):
                p.fill_(1.0)                       # LayerScale 1 so the blocks are not degenerate
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif p.dim() >= 2 and "pos_embed" not in name and "token" not in name:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))

    def interpolated_pos_embed(self, gh, gw):
        """Upstream interpolate_pos_encoding: bicubic, scale_factor=(g+0.1)/37, antialias off."""
        pe = self.pos_embed.float()
        n = pe.shape[1] - 1
        m = int(math.sqrt(n))
        if gh * gw == n and gh == gw:
            return pe
        patch = pe[:, 1:].reshape(1, m, m, self.dim).permute(0, 3, 1, 2)
        patch = F.interpolate(patch, scale_factor=((gh + 0.1) / m, (gw + 0.1) / m), mode="bicubic")
        assert patch.shape[-2:] == (gh, gw)
        patch = patch.permute(0, 2, 3, 1).reshape(1, gh * gw, self.dim)
        return torch.cat([pe[:, :1], patch], dim=1)

    @torch.no_grad()
    def forward_features(self, x):
        B, _, H, W = x.shape
        gh, gw = H // self.patch_size, W // self.patch_size
        tok = self.patch_embed.proj(x).flatten(2).transpose(1, 2)
        tok = torch.cat([self.cls_token.expand(B, -1, -1), tok], dim=1) + self.interpolated_pos_embed(gh, gw)
        for blk in self.blocks:
            tok = blk(tok)
        return {"x_prenorm": tok, "x_norm_patchtokens": self.norm(tok)[:, 1:]}


@torch.no_grad()
def ae_features(vit, images, chunk=64):
    """ae_net.py:55-69: pre-norm patch tokens, CLS dropped, b (h w) c -> b c h w, L2 normalised over c."""
    outs = []
    for i in range(0, images.shape[0], chunk):
        t = vit.forward_features(images[i:i + chunk])["x_prenorm"][:, 1:, :]
        b, n, c = t.shape
        g = int(math.sqrt(n))
        outs.append(t.reshape(b, g, g, c).permute(0, 3, 1, 2))
    return F.normalize(torch.cat(outs, 0), dim=1)


# ------------------------------------------------------------------------------------------------------------
# a6  IST ResNet backbone   (resnet.py:26-50,318-381)
# ------------------------------------------------------------------------------------------------------------
class _BB(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn1, self.bn2 = nn.BatchNorm2d(cout), nn.BatchNorm2d(cout)
        self.downsample = None if stride == 1 else nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                                                 nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)))))
        if self.downsample is not None:
            x = self.downsample(x)
        return F.relu(x + y)


class ISTBackbonePort(nn.Module):
    def __init__(self, initial_dim=128, block_dims=(128, 192, 256, 512), descriptor_size=256, input_size=256,
                 seed: Optional[int] = 8):
        super().__init__()
        self.input_size = input_size
        self.conv1 = nn.Conv2d(3, initial_dim, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(initial_dim)
        dims, cin, layers = list(block_dims), initial_dim, []
        for i, d in enumerate(dims):
            layers.append(nn.Sequential(_BB(cin, d, 1 if i == 0 else 2), _BB(d, d, 1)))
            cin = d
        self.layer1, self.layer2, self.layer3, self.layer4 = layers
        self.layer4_outconv = nn.Conv2d(dims[3], descriptor_size, 1, bias=False)
        if seed is not None:
            g = torch.Generator().manual_seed(seed)
            with torch.no_grad():
                for name, p in self.named_parameters():
                    if p.dim() == 4:
                        p.copy_(torch.randn(p.shape, generator=g) * math.sqrt(2.0 / p[0].numel()))
                    elif name.endswith("weight"):
                        p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                    else:
                        p.copy_(0.05 * torch.randn(p.shape, generator=g))
                for name, b in self.named_buffers():
                    if name.endswith("running_mean"):
                        b.copy_(0.1 * torch.randn(b.shape, generator=g))
                    elif name.endswith("running_var"):
                        b.copy_(1.0 + 0.2 * torch.rand(b.shape, generator=g))
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        x = F.interpolate(x, (self.input_size, self.input_size), mode="bilinear", align_corners=True)
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.layer4_outconv(x)


# ------------------------------------------------------------------------------------------------------------
# a3+a4+a5+a7+a8+a9  eval_retrieval sequencing   (gigaPose.py:497-604)
# ------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def crop_resize_pad(xyxy_boxes, images, target_size=224):
    """Row f3: CPU restatement of `CropResizePad.__call__` (reference src/utils/crop.py:16-61) as explicit index maps.

    xyxy_boxes [n,4] (any integer/float dtype; truncated to int64 as `BoundingBox.convert_long`, bbox.py:18-22),
    images [n,C,H,W] -> dict(M [n,3,3] f32, images [n,C,target,target]).  Per detection (crop.py:22-55):
      crop image[:, y1:y2, x1:x2] (python slicing: upper bounds clip to the image);
      nearest resize by scale = target / max(w_box, h_box) (a float32 tensor value, passed on as a python float):
        out size floor(size * scale) in double, source index min(floor(dst * float32(1 / scale)), size - 1)
        (ATen's small-output kernel special-cases unchanged / doubled sizes, see `nearest_map`);
      if the resized crop is not square: centred zero padding to target x target;
      nearest resize to (target, target) by size ratio (float32(in / out)): the identity unless a pixel went missing;
      M = M_resize_pad @ M_crop.
    """
    boxes = torch.as_tensor(xyxy_boxes).long()
    n, C, H, W = images.shape
    T = int(target_size)
    out = torch.zeros(n, C, T, T, dtype=images.dtype)
    Ms = torch.zeros(n, 3, 3, dtype=torch.float32)
    sizes = torch.stack([boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]], dim=-1)
    scales = T / torch.max(sizes, dim=-1)[0]                               # float32 tensor (crop.py:20)
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)

    def nearest_map(out_size, in_size, inv, small):
        # ATen CPU nearest (UpSampleKernel.cpp): source index min(floorf(dst * inv), in - 1) in float32.  Outputs with
        # out_h + out_w <= 128 go through a second kernel (`_use_vectorized_kernel_cond_2d`) whose `nearest_idx`
        # special-cases an unchanged size (identity) and an exactly doubled size (dst >> 1); at the shipped target size
        # 224 that kernel is never selected.
        dst = torch.arange(out_size)
        if small and out_size == in_size:
            return dst
        if small and out_size == 2 * in_size:
            return dst >> 1
        idx = torch.floor(dst.to(torch.float32) * inv).long()
        return torch.clamp(idx, max=in_size - 1)

    for i in range(n):
        x1, y1, x2, y2 = (int(v) for v in boxes[i])
        assert 0 <= x1 < x2 and 0 <= y1 < y2, "boxes need a non-negative top-left corner and positive size"
        scale = scales[i].item()
        ch, cw = min(y2, H) - y1, min(x2, W) - x1                          # crop size after slicing
        rh, rw = int(math.floor(float(ch) * scale)), int(math.floor(float(cw) * scale))
        inv = f32(1.0 / scale)
        small = rh + rw <= 128
        rows, cols = y1 + nearest_map(rh, ch, inv, small), x1 + nearest_map(rw, cw, inv, small)
        pad_left = pad_top = 0
        ph, pw = rh, rw
        if rw / rh != 1:                                                   # crop.py:37-46
            pad_top = (T - rh) // 2
            pad_bottom = max(T - rh - pad_top, 0)
            pad_left = max((T - rw) // 2, 0)
            pad_right = T - rw - pad_left
            ph, pw = rh + pad_top + pad_bottom, rw + pad_left + pad_right
        # final resize to (T, T): padded row / column of every output pixel, then back to crop coordinates
        prow = nearest_map(T, ph, f32(ph) / f32(T), 2 * T <= 128) - pad_top
        pcol = nearest_map(T, pw, f32(pw) / f32(T), 2 * T <= 128) - pad_left
        ok_r, ok_c = (prow >= 0) & (prow < rh), (pcol >= 0) & (pcol < rw)
        src_r, src_c = rows[prow.clamp(0, rh - 1)], cols[pcol.clamp(0, rw - 1)]
        patch = images[i][:, src_r][:, :, src_c]
        out[i] = patch * (ok_r[:, None] & ok_c[None, :]).to(images.dtype)
        M_crop, M_rp = torch.eye(3), torch.eye(3)
        M_crop[:2, 2] = -boxes[i, :2].float()
        M_rp[:2, :2] *= scales[i]
        if rw / rh != 1:
            M_rp[:2, 2] = torch.tensor([pad_left, pad_top], dtype=torch.float32)
        Ms[i] = torch.matmul(M_rp, M_crop)
    return {"M": Ms, "images": out}


def retrieval(ref_inputs, regressor, k=5, sim_threshold=0.5, patch_threshold=3, sub_batch=None):
    """`ref_inputs` = gigapose_b200.synth.to_reference_layout(case): features already extracted (feature-level)."""
    B = ref_inputs["tar_feat"].shape[0]
    sub = sub_batch or B
    parts = []
    for b0 in range(0, B, sub):                                                    # gigaPose.py:500-536
        sl = slice(b0, b0 + sub)
        parts.append(similarity_search(ref_inputs["src_feats"][sl], ref_inputs["tar_feat"][sl],
                                       ref_inputs["src_masks"][sl], ref_inputs["tar_mask"][sl],
                                       k, sim_threshold, patch_threshold))
    pred = {n: torch.cat([p[n] for p in parts], 0) for n in parts[0]}
    rel_scale = torch.zeros(B, k, G * G)
    rel_inpl = torch.zeros(B, k, G * G, 2)
    bi = torch.arange(B)
    for kk in range(k):                                                            # gigaPose.py:545-575
        src_ist = ref_inputs["src_ist"][bi, pred["id_src"][:, kk]]
        rel_scale[:, kk], rel_inpl[:, kk] = ist_mlp(regressor, src_ist, ref_inputs["tar_ist"],
                                                    pred["src_pts"][:, kk], pred["tar_pts"][:, kk])
    pred["relScale"], pred["relInplane"] = rel_scale, rel_inpl
    M, failed, in_src, in_tar, in_sc = ransac(pred["src_pts"], pred["tar_pts"], rel_scale, rel_inpl)
    pred.update(M=M, idx_failed=failed, ransac_src_pts=in_src, ransac_tar_pts=in_tar, ransac_scores=in_sc)
    scores = torch.sum(in_sc, dim=2) / (G * G)                                    # gigaPose.py:588
    pred["scores"] = scores
    order = torch.argsort(scores, dim=1, descending=True)                         # :590-595
    for n, v in list(pred.items()):
        pred[n] = v[bi[:, None], order]
    pred["pred_poses"] = pose_recovery(ref_inputs["tar_label"], ref_inputs["tar_K"], ref_inputs["tar_M"],
                                       pred["id_src"], pred["M"].clone(), ref_inputs["template_K"],
                                       ref_inputs["template_Ms"], ref_inputs["template_poses"])
    return pred

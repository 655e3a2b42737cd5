"""Planted synthetic workloads for the GigaPose hot path.

Random unit vectors are useless as a workload (cosine of random 1024-d vectors ~ N(0, 1/32) << 0.5, so
every patch is thresholded away).  The generator plants structure instead:

* every object ``o`` owns a base descriptor field ``F_o[256, 1024]``;
* template ``tau`` of that object is ``normalise(warp_tau(F_o) + sigma_tau * noise)`` where ``warp_tau`` is a
  patch-grid resampling under a known 2-D similarity (scale, in-plane angle, shift) and ``sigma_tau`` orders
  the templates by quality (one planted best view ``tau*`` per object);
* a query of object ``o`` is ``normalise(F_o + 0.3 * noise)``.

Everything is produced in the *kernel-native* layout (patch-major ``[.., 256, 1024]`` features, 16x16 masks);
``to_reference_layout`` converts a (small) case into the tensors the reference modules take
(``[b, T, 1024, 16, 16]`` features, 224x224 masks).  Pure torch, deterministic per seed, device agnostic.
This file is product-side (bench.py and the tests use it); it does not depend on ``oracle/``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

GRID = 16          # patches per side (224 / 14)
P = GRID * GRID    # 256 patches
C_AE = 1024        # DINOv2 ViT-L/14 descriptor size
C_IST = 256        # IST descriptor size

# LINEMOD intrinsics (public BOP camera.json values), used for query and template cameras alike
LM_K = ((572.4114, 0.0, 325.2611), (0.0, 573.57043, 242.04899), (0.0, 0.0, 1.0))


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def fibonacci_view_poses(T: int, distance: float = 400.0, dtype=torch.float32) -> torch.Tensor:
    """T object-to-camera poses looking at the origin from a Fibonacci sphere, translation (0,0,distance)."""
    i = torch.arange(T, dtype=torch.float64) + 0.5
    phi = torch.acos(1 - 2 * i / T)
    theta = math.pi * (1 + 5 ** 0.5) * i
    cam = torch.stack([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], 1)
    z = -cam                                               # camera looks at the origin
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64).expand(T, 3).clone()
    degenerate = (torch.cross(up, z, dim=1).norm(dim=1) < 1e-6)
    up[degenerate] = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    x = F.normalize(torch.cross(up, z, dim=1), dim=1)
    y = torch.cross(z, x, dim=1)
    R = torch.stack([x, y, z], dim=1)                      # rows = camera axes in object frame
    pose = torch.eye(4, dtype=torch.float64).repeat(T, 1, 1)
    pose[:, :3, :3] = R
    pose[:, 2, 3] = distance
    return pose.to(dtype)


def disc_mask(radius: torch.Tensor, cx: torch.Tensor, cy: torch.Tensor) -> torch.Tensor:
    """Binary 16x16 disc masks, flattened to [..., 256]."""
    dev = radius.device
    ys, xs = torch.meshgrid(torch.arange(GRID, device=dev), torch.arange(GRID, device=dev), indexing="ij")
    xs = xs.reshape(-1).float()
    ys = ys.reshape(-1).float()
    d2 = (xs - cx[..., None]) ** 2 + (ys - cy[..., None]) ** 2
    return (d2 <= (radius[..., None] ** 2)).float()


@dataclass
class FeatureCase:
    """Feature-level workload: inputs of rows a3-a9 of SURVEY.md §8."""
    B: int
    O: int
    T: int
    seed: int
    bank_feat: torch.Tensor      # [O, T, 256, 1024] f32, unit norm over C (patch-major)
    bank_mask16: torch.Tensor    # [O, T, 256] f32
    bank_ist: torch.Tensor       # [O, T, 256(c), 16, 16] f32
    bank_M: torch.Tensor         # [O, T, 3, 3]
    bank_poses: torch.Tensor     # [O, T, 4, 4]
    bank_K: torch.Tensor         # [O, 3, 3]
    q_feat: torch.Tensor         # [B, 256, 1024] f32 unit norm
    q_mask16: torch.Tensor       # [B, 256]
    q_ist: torch.Tensor          # [B, 256(c), 16, 16]
    q_label: torch.Tensor        # [B] int64, 1-based object id (reference convention, gigaPose.py:514-521)
    q_K: torch.Tensor            # [B, 3, 3]
    q_M: torch.Tensor            # [B, 3, 3]
    planted: Dict[str, torch.Tensor] = field(default_factory=dict)

    def to(self, device) -> "FeatureCase":
        kw = {}
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                kw[k] = v.to(device)
            elif isinstance(v, dict):
                kw[k] = {kk: vv.to(device) if torch.is_tensor(vv) else vv for kk, vv in v.items()}
            else:
                kw[k] = v
        return FeatureCase(**kw)


def _warp_index(scale, angle, shift):
    """For every template patch s: index of the base-field patch it shows (or -1 when outside the grid).

    Template position p_s = c + d + s * R(angle) (q - c)  =>  q = c + R(-angle) (p_s - c - d) / s.
    scale/angle: [...]; shift: [..., 2].  Returns int64 [..., 256].
    """
    dev = scale.device
    ys, xs = torch.meshgrid(torch.arange(GRID, device=dev), torch.arange(GRID, device=dev), indexing="ij")
    c = (GRID - 1) / 2.0
    px = xs.reshape(-1).float() - c - shift[..., 0:1]
    py = ys.reshape(-1).float() - c - shift[..., 1:2]
    ca, sa = torch.cos(angle)[..., None], torch.sin(angle)[..., None]
    qx = (ca * px + sa * py) / scale[..., None] + c
    qy = (-sa * px + ca * py) / scale[..., None] + c
    qxr, qyr = torch.round(qx).long(), torch.round(qy).long()
    ok = (qxr >= 0) & (qxr < GRID) & (qyr >= 0) & (qyr < GRID)
    idx = qyr * GRID + qxr
    return torch.where(ok, idx, torch.full_like(idx, -1))


def make_feature_case(B: int, O: int, T: int, seed: int = 42, device="cpu",
                      q_noise: float = 0.3, sigma_lo: float = 0.3, sigma_hi: float = 1.2,
                      labels: Optional[torch.Tensor] = None, obj_chunk: int = 4) -> FeatureCase:
    dev = torch.device(device)
    g = _gen(seed, dev)
    gc = _gen(seed + 7, "cpu")

    # --- per (object, template) planted similarity + quality ordering
    scale = torch.empty(O, T).uniform_(0.8, 1.25, generator=gc)
    angle = torch.empty(O, T).uniform_(-math.pi, math.pi, generator=gc)
    shift = torch.randint(-2, 3, (O, T, 2), generator=gc).float()
    rank = torch.stack([torch.randperm(T, generator=gc) for _ in range(O)])          # rank 0 = planted best view
    sigma = sigma_lo + (sigma_hi - sigma_lo) * rank.float() / max(T - 1, 1)
    t_rad = torch.randint(5, 8, (O, T), generator=gc).float()
    t_cx = (GRID - 1) / 2.0 + torch.randint(-1, 2, (O, T), generator=gc).float() * 0.5
    t_cy = (GRID - 1) / 2.0 + torch.randint(-1, 2, (O, T), generator=gc).float() * 0.5
    bank_mask16 = disc_mask(t_rad, t_cx, t_cy)
    widx = _warp_index(scale, angle, shift)                                             # [O, T, 256]

    bank_feat = torch.empty(O, T, P, C_AE, device=dev)
    base = torch.empty(O, P, C_AE, device=dev)
    for o0 in range(0, O, obj_chunk):
        o1 = min(O, o0 + obj_chunk)
        n = o1 - o0
        Fo = torch.randn(n, P, C_AE, generator=g, device=dev)
        base[o0:o1] = Fo
        w = widx[o0:o1].to(dev)
        gathered = torch.gather(Fo[:, None].expand(n, T, P, C_AE), 2,
                                w.clamp(min=0)[..., None].expand(n, T, P, C_AE))
        noise = torch.randn(n, T, P, C_AE, generator=g, device=dev)
        outside = (w < 0)[..., None]
        feat = torch.where(outside, noise, gathered + sigma[o0:o1].to(dev)[..., None, None] * noise)
        bank_feat[o0:o1] = F.normalize(feat, dim=-1)
        del gathered, noise, feat

    bank_ist = torch.randn(O, T, C_IST, GRID, GRID, generator=g, device=dev)

    # --- queries
    if labels is None:
        labels = torch.randint(1, O + 1, (B,), generator=gc)
    labels = labels.long()
    q_feat = F.normalize(base[(labels - 1).to(dev)] + q_noise * torch.randn(B, P, C_AE, generator=g, device=dev), dim=-1)
    q_rad = torch.randint(5, 8, (B,), generator=gc).float()
    q_mask16 = disc_mask(q_rad, torch.full((B,), (GRID - 1) / 2.0), torch.full((B,), (GRID - 1) / 2.0))
    q_ist = torch.randn(B, C_IST, GRID, GRID, generator=g, device=dev)

    # --- cameras, crop matrices, template poses
    K = torch.tensor(LM_K)
    bank_K = K.repeat(O, 1, 1)
    q_K = K.repeat(B, 1, 1)

    def crop_M(n):
        s = torch.empty(n).uniform_(0.6, 1.8, generator=gc)
        cx = torch.empty(n).uniform_(150, 490, generator=gc)
        cy = torch.empty(n).uniform_(120, 360, generator=gc)
        M = torch.zeros(n, 3, 3)
        M[:, 0, 0] = s
        M[:, 1, 1] = s
        M[:, 0, 2] = 112.0 - s * cx
        M[:, 1, 2] = 112.0 - s * cy
        M[:, 2, 2] = 1
        return M

    bank_M = crop_M(O * T).reshape(O, T, 3, 3)
    q_M = crop_M(B)
    bank_poses = fibonacci_view_poses(T).repeat(O, 1, 1, 1)

    planted = dict(scale=scale, angle=angle, shift=shift, rank=rank, sigma=sigma,
                   best_template=torch.argmin(rank, dim=1), warp_index=widx)
    to = lambda x: x.to(dev)
    return FeatureCase(B=B, O=O, T=T, seed=seed,
                       bank_feat=bank_feat, bank_mask16=to(bank_mask16), bank_ist=bank_ist,
                       bank_M=to(bank_M), bank_poses=to(bank_poses), bank_K=to(bank_K),
                       q_feat=q_feat, q_mask16=to(q_mask16), q_ist=q_ist, q_label=to(labels),
                       q_K=to(q_K), q_M=to(q_M), planted=planted)


def mask16_to_224(mask16: torch.Tensor) -> torch.Tensor:
    """[..., 256] -> [..., 224, 224] by x14 replication (nearest 224->16 sampling then recovers mask16)."""
    m = mask16.reshape(*mask16.shape[:-1], GRID, GRID)
    return m.repeat_interleave(14, dim=-2).repeat_interleave(14, dim=-1)


def to_reference_layout(case: FeatureCase) -> Dict[str, torch.Tensor]:
    """Tensors in the layout the reference modules consume (gigaPose.py:513-531).  Small cases only."""
    lab = case.q_label - 1
    src_feats = case.bank_feat[lab]                                   # [B, T, 256, 1024]
    B, T = src_feats.shape[:2]
    src_feats = src_feats.permute(0, 1, 3, 2).reshape(B, T, C_AE, GRID, GRID).contiguous()
    tar_feat = case.q_feat.permute(0, 2, 1).reshape(B, C_AE, GRID, GRID).contiguous()
    return dict(
        src_feats=src_feats, tar_feat=tar_feat,
        src_masks=mask16_to_224(case.bank_mask16[lab]), tar_mask=mask16_to_224(case.q_mask16),
        src_ist=case.bank_ist[lab], tar_ist=case.q_ist,
        tar_label=case.q_label, tar_K=case.q_K, tar_M=case.q_M,
        template_K=case.bank_K, template_Ms=case.bank_M, template_poses=case.bank_poses,
    )


# ------------------------------------------------------------------------------------------------------------
# crop-level workload (adds rows a1/a6): smooth random textures inside a disc, CLIP-normalised
# ------------------------------------------------------------------------------------------------------------
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def make_crops(n: int, seed: int, device="cpu", radius_px: float = 90.0):
    """n synthetic 224x224 crops: low-pass random texture inside a centred disc, zero background,
    CLIP mean/std normalised (configs/data/transform.yaml:5-7).  Returns (rgb [n,3,224,224], mask [n,224,224])."""
    dev = torch.device(device)
    g = _gen(seed, dev)
    low = torch.randn(n, 3, 28, 28, generator=g, device=dev)
    tex = F.interpolate(low, size=(224, 224), mode="bilinear", align_corners=False)
    tex = (0.5 + 0.25 * tex).clamp(0, 1)
    ys, xs = torch.meshgrid(torch.arange(224, device=dev), torch.arange(224, device=dev), indexing="ij")
    mask = (((xs - 111.5) ** 2 + (ys - 111.5) ** 2) <= radius_px ** 2).float()
    mask = mask.expand(n, 224, 224).contiguous()
    rgb = tex * mask[:, None]
    mean = torch.tensor(CLIP_MEAN, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, device=dev).view(1, 3, 1, 1)
    return (rgb - mean) / std, mask

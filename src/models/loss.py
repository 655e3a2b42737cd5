"""Training losses named by configs/model/large.yaml:11-24.  Training is outside the B200 hot path; the classes
exist because Hydra instantiates them even for `test.py` (reference `src/models/loss.py`)."""
import torch
import torch.nn.functional as F
from torch import nn


def cosine_similarity(a, b, normalize=True):
    sim = a @ b.t()
    if normalize:
        sim = sim / (a.norm(dim=1, keepdim=True) * b.norm(dim=1, keepdim=True).t()).clamp(min=1e-8)
    return sim


class InfoNCE(nn.Module):
    def __init__(self, tau=0.1):
        super().__init__()
        self.tau = tau

    def forward(self, query_feat, ref_feats, labels):
        logits = F.normalize(query_feat, dim=1) @ F.normalize(ref_feats, dim=1).t()
        return F.cross_entropy(logits / self.tau, labels)


class ScaleLoss(nn.Module):
    def __init__(self, loss="l2", log=False):
        super().__init__()
        self.loss, self.log = loss, log

    def forward(self, pred_scale, gt_scale):
        if self.log:
            pred_scale, gt_scale = torch.log(pred_scale.clamp(min=1e-6)), torch.log(gt_scale)
        return F.l1_loss(pred_scale, gt_scale) if self.loss == "l1" else F.mse_loss(pred_scale, gt_scale)


class InplaneLoss(nn.Module):
    def __init__(self, loss="l2", normalize=False):
        super().__init__()
        self.loss, self.normalize = loss, normalize

    def forward(self, pred_cos_sin, gt_cos_sin):
        if self.normalize:
            pred_cos_sin = F.normalize(pred_cos_sin, dim=1)
        if self.loss == "geodesic":
            cos = (pred_cos_sin * gt_cos_sin).sum(dim=1).clamp(-1, 1)
            return torch.acos(cos).mean()
        return F.l1_loss(pred_cos_sin, gt_cos_sin) if self.loss == "l1" else F.mse_loss(pred_cos_sin, gt_cos_sin)

"""Shared helpers for the GPU parity tests: load a synthetic FeatureCase into an Engine, run the oracle."""
from __future__ import annotations

import torch

from gigapose_b200 import synth
from gigapose_b200.engine import Engine

INT_KEYS = ["id_src", "tar_pts", "src_pts", "idx_failed", "ransac_scores", "ransac_src_pts", "ransac_tar_pts"]
FLOAT_KEYS = ["score_src", "score_pts", "relScale", "relInplane", "M", "scores", "pred_poses"]


def engine_from_case(case: synth.FeatureCase, device="cuda:0", precision="fp32_split", regressor=None,
                     shard_rank=0, shard_world=1, max_batch=None) -> Engine:
    """Loads the (already unit-norm, patch-major) synthetic bank; the kernel applies the matching-time
    normalisation (matching.py:229), i.e. norm_passes=1, exactly like the oracle does on the same tensors."""
    local = list(range(shard_rank, case.T, shard_world))
    eng = Engine(case.O, len(local), max_batch or case.B, device=device, precision=precision,
                 shard_rank=shard_rank, shard_world=shard_world, num_templates_global=case.T)
    sel = torch.tensor(local)
    for o in range(case.O):
        eng.bank_write(o, 0, case.bank_feat[o, sel], case.bank_mask16[o, sel].reshape(-1, 16, 16),
                       ist_feat=case.bank_ist[o, sel], norm_passes=1)
    eng.set_poses(case.bank_K, case.bank_M, case.bank_poses)
    if regressor is not None:
        eng.set_ist_weights(regressor)
    return eng


def run_engine(eng: Engine, case: synth.FeatureCase):
    return eng.retrieve(case.q_feat, case.q_mask16.reshape(-1, 16, 16), case.q_label - 1, case.q_ist, case.q_K,
                        case.q_M, norm_passes=1)


def cpu(d):
    return {k: v.detach().cpu() for k, v in d.items()}

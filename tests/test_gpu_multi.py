"""-m gpu: the sharded similarity search.  Two template-interleaved Engine shards (both on cuda:0, standing in
for two ranks; the collective itself is covered by the gloo test and by bench.py --gpus N) must merge to exactly
the single-bank result, through the packed candidate buffer + rank-stride merge kernel used on multi-GPU runs."""
import pytest
import torch

from gigapose_b200 import multigpu, synth
from oracle import port

from helpers import cpu, engine_from_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,G", [(12, 2), (13, 3)])
def test_two_shards_merge_to_single_bank_result(T, G):
    case = synth.make_feature_case(B=5, O=2, T=T, seed=23)
    reg = port.RegressorPort(seed=3)
    ref = port.retrieval(synth.to_reference_layout(case), reg)
    B, k = case.B, 5
    dev = torch.device("cuda:0")
    gathered, total = multigpu.alloc_packed(B, k, dev, world=G)
    engines = []
    for r in range(G):
        eng = engine_from_case(case, regressor=reg, shard_rank=r, shard_world=G)
        eng.set_queries(case.q_feat, case.q_mask16.reshape(-1, 16, 16), case.q_label - 1)
        mine = multigpu.field_views(gathered, B, k, total, rank_slot=r)
        eng.sim_candidates(out=mine)
        local_m = eng.topk_merge(dict(mine, rel_scale=None, rel_inplane=None), G=1)
        rs, ri = eng.ist_mlp(case.q_ist, local_m)
        mine["rel_scale"].copy_(rs)
        mine["rel_inplane"].copy_(ri)
        engines.append(eng)
    eng = engines[0]
    g0 = multigpu.field_views(gathered, B, k, total, rank_slot=0)
    m, rel_scale, rel_inplane = eng.topk_merge(g0, G=G, rank_stride_bytes=total)
    rr = eng.ransac(m, rel_scale, rel_inplane)
    out = cpu(eng.sort_and_pose(case.q_K, case.q_M, m, rel_scale, rel_inplane, rr))
    for key in ("id_src", "tar_pts", "src_pts", "ransac_scores", "ransac_src_pts"):
        assert torch.equal(out[key], ref[key]), key
    assert torch.equal(out["idx_failed"], ref["idx_failed"])
    assert torch.allclose(out["relScale"], ref["relScale"], atol=2e-5, rtol=1e-5)
    err = (out["pred_poses"] - ref["pred_poses"]).abs()
    err[..., :3, 3] /= ref["pred_poses"][..., :3, 3].abs().clamp(min=1.0)
    assert float(err.max()) < 1e-3

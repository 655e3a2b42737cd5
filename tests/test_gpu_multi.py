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
    assert torch.allclose(out["relScale"], ref["relScale"], atol=1e-4, rtol=1e-5)
    err = (out["pred_poses"] - ref["pred_poses"]).abs()
    err[..., :3, 3] /= ref["pred_poses"][..., :3, 3].abs().clamp(min=1.0)
    assert float(err.max()) < 1e-3


@pytest.mark.parametrize("T,G,B", [(12, 2, 6), (13, 3, 5)])
def test_query_sharded_tail_with_replicated_ist_bank(T, G, B):
    """The round-2 pipeline on one device: G descriptor shards with the IST bank replicated (global ids), light candidate
    records merged across shards, rows a5-a9 computed per detection WINDOW (uneven windows included) -- every window
    equals the oracle's rows."""
    from gigapose_b200.engine import Engine
    case = synth.make_feature_case(B=B, O=2, T=T, seed=29)
    reg = port.RegressorPort(seed=4)
    ref = port.retrieval(synth.to_reference_layout(case), reg)
    k = 5
    dev = torch.device("cuda:0")
    packed, total = multigpu.alloc_packed(B, k, dev, world=G, light=True)
    engines = []
    for r in range(G):
        ids = multigpu.shard_template_ids(T, r, G)
        eng = Engine(case.O, len(ids), B, device=dev, shard_rank=r, shard_world=G, num_templates_global=T,
                     ist_bank_global=True)
        sel = torch.tensor(ids)
        for o in range(case.O):
            eng.bank_write(o, 0, case.bank_feat[o, sel], case.bank_mask16[o, sel].reshape(-1, 16, 16), norm_passes=1)
            eng.bank_write_ist(o, 0, case.bank_ist[o])                  # all T templates, global ids
        eng.set_poses(case.bank_K, case.bank_M, case.bank_poses)
        eng.set_ist_weights(reg)
        eng.set_queries(case.q_feat, case.q_mask16.reshape(-1, 16, 16), case.q_label - 1)
        eng.sim_candidates(out=multigpu.field_views(packed, B, k, total, rank_slot=r, light=True))
        engines.append(eng)
    slot0 = multigpu.field_views(packed, B, k, total, rank_slot=0, light=True)
    for r, eng in enumerate(engines):
        m = eng.topk_merge(dict(slot0, rel_scale=None, rel_inplane=None), G=G, rank_stride_bytes=total)
        lo, hi = multigpu.window(B, r, G)
        if hi == lo:
            continue
        mw = {kk: v[lo:hi] for kk, v in m.items()}
        rs, ri = eng.ist_mlp(case.q_ist[lo:hi], mw, b0=lo)
        rr = eng.ransac(mw, rs, ri)
        out = cpu(eng.sort_and_pose(case.q_K[lo:hi], case.q_M[lo:hi], mw, rs, ri, rr, b0=lo))
        for key in ("id_src", "tar_pts", "src_pts", "ransac_scores", "ransac_src_pts", "idx_failed"):
            assert torch.equal(out[key], ref[key][lo:hi]), (r, key)
        assert torch.allclose(out["relScale"], ref["relScale"][lo:hi], atol=1e-4, rtol=1e-5)
        err = (out["pred_poses"] - ref["pred_poses"][lo:hi]).abs()
        err[..., :3, 3] /= ref["pred_poses"][lo:hi][..., :3, 3].abs().clamp(min=1.0)
        assert float(err.max()) < 1e-3


def test_sort_pred_by_inliers_false_keeps_retrieval_order():
    """gigaPose.py:590: with sort_pred_by_inliers=False the k hypotheses stay in similarity order."""
    case = synth.make_feature_case(B=4, O=2, T=10, seed=31)
    reg = port.RegressorPort(seed=5)
    eng = engine_from_case(case, regressor=reg)
    eng.set_queries(case.q_feat, case.q_mask16.reshape(-1, 16, 16), case.q_label - 1)
    m = eng.sim_topk()
    rs, ri = eng.ist_mlp(case.q_ist, m)
    rr = eng.ransac(m, rs, ri)
    keep = cpu(eng.sort_and_pose(case.q_K, case.q_M, m, rs, ri, rr, sort_by_inliers=False))
    srt = cpu(eng.sort_and_pose(case.q_K, case.q_M, m, rs, ri, rr, sort_by_inliers=True))
    assert torch.equal(keep["id_src"], m["id_src"].cpu())
    assert torch.equal(keep["scores"], rr["inlier_count"].cpu().float() / 256)
    order = torch.argsort(keep["scores"], dim=1, descending=True, stable=True)
    assert torch.equal(torch.gather(keep["id_src"], 1, order), srt["id_src"])
    assert torch.equal(torch.gather(keep["scores"], 1, order), srt["scores"])

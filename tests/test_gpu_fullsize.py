"""-m gpu: BASELINE.json configs[1] at FULL size (8 objects x 162 templates, batch 32), checked through
size-independent properties (the CPU oracle would need minutes at this size): determinism, permutation equivariance,
shard-merge equivalence, structural invariants of every output, and the planted structure of the generator."""
import pytest
import torch

from gigapose_b200 import multigpu, synth
from oracle import port

from helpers import cpu, engine_from_case, run_engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    case = synth.make_feature_case(B=32, O=8, T=162, seed=42)
    reg = port.RegressorPort(seed=9)
    eng = engine_from_case(case, regressor=reg)
    out = cpu(run_engine(eng, case))
    return case, reg, eng, out


def test_deterministic_bits(c2):
    case, reg, eng, out = c2
    again = cpu(run_engine(eng, case))
    for k, v in out.items():
        assert torch.equal(v, again[k]), k


def test_structural_invariants(c2):
    case, reg, eng, out = c2
    B, K = out["id_src"].shape
    assert (out["id_src"] >= 0).all() and (out["id_src"] < case.T).all()
    for b in range(B):
        assert len(set(out["id_src"][b].tolist())) == K                  # k distinct templates per detection
    s = out["scores"]
    assert (s[:, :-1] >= s[:, 1:]).all()                                 # re-sorted by inliers (gigaPose.py:590-595)
    valid = out["src_pts"][..., 0] != -1
    assert torch.equal(valid, out["tar_pts"][..., 0] != -1)
    assert ((out["relScale"] == -1000) == ~valid).all()                  # ist_net.py:110-113
    n_valid = valid.sum(-1)
    n_inl = out["ransac_scores"].sum(-1)
    assert (n_inl <= (n_valid - 1).clamp(min=0)).all()                   # proposer excluded (ransac.py:29-33)
    assert torch.allclose(s, n_inl.float() / 256)
    assert torch.equal(out["idx_failed"], (n_inl == 0) & (n_valid > 0))
    # inlier lists are compacted: first n_inl slots valid, rest -1
    slots = torch.arange(256)[None, None]
    assert torch.equal(out["ransac_src_pts"][..., 0] != -1, slots < n_inl[..., None])
    # tar_pts of a valid slot is the slot's own patch coordinate (format_prediction, matching.py:29-61)
    t = torch.arange(256)
    assert torch.equal(out["tar_pts"][..., 0][valid], (t % 16).expand(B, K, 256)[valid])
    assert torch.equal(out["tar_pts"][..., 1][valid], (t // 16).expand(B, K, 256)[valid])
    # rotations are orthonormal, last row is (0,0,0,1)
    R = out["pred_poses"][..., :3, :3]
    ok = ~out["idx_failed"]
    eye = torch.eye(3).expand(B, K, 3, 3)
    assert torch.allclose((R @ R.transpose(-1, -2))[ok], eye[ok], atol=1e-4)
    assert torch.equal(out["pred_poses"][..., 3, :], torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(B, K, 4))


def test_permutation_equivariance(c2):
    """Shuffling the batch shuffles the results and nothing else (queries only interact through the bank)."""
    case, reg, eng, out = c2
    perm = torch.randperm(case.B, generator=torch.Generator().manual_seed(1))
    out_p = cpu(eng.retrieve(case.q_feat[perm], case.q_mask16[perm].reshape(-1, 16, 16), case.q_label[perm] - 1,
                             case.q_ist[perm], case.q_K[perm], case.q_M[perm]))
    for k in ("id_src", "src_pts", "tar_pts", "ransac_scores", "scores", "pred_poses", "M"):
        assert torch.equal(out_p[k], out[k][perm]), k


def test_two_shards_equal_one_bank_at_full_size(c2):
    case, reg, eng, out = c2
    B, k, G = case.B, 5, 2
    dev = eng.device
    gathered, total = multigpu.alloc_packed(B, k, dev, world=G)
    shard_engs = []
    for r in range(G):
        e = engine_from_case(case, regressor=reg, shard_rank=r, shard_world=G)
        e.set_queries(case.q_feat, case.q_mask16.reshape(-1, 16, 16), case.q_label - 1)
        mine = multigpu.field_views(gathered, B, k, total, rank_slot=r)
        e.sim_candidates(out=mine)
        lm = e.topk_merge(dict(mine, rel_scale=None, rel_inplane=None), G=1)
        rs, ri = e.ist_mlp(case.q_ist, lm)
        mine["rel_scale"].copy_(rs)
        mine["rel_inplane"].copy_(ri)
        shard_engs.append(e)
    e0 = shard_engs[0]
    m, rs, ri = e0.topk_merge(multigpu.field_views(gathered, B, k, total, 0), G=G, rank_stride_bytes=total)
    rr = e0.ransac(m, rs, ri)
    sharded = cpu(e0.sort_and_pose(case.q_K, case.q_M, m, rs, ri, rr))
    for key in ("id_src", "src_pts", "tar_pts", "ransac_scores", "scores", "relScale", "M", "pred_poses"):
        assert torch.equal(sharded[key], out[key]), key


def test_first_stage_matches_oracle_on_a_slice(c2):
    """The oracle on 4 of the 32 queries (full 162-template banks of their objects)."""
    case, reg, eng, out = c2
    sel = torch.tensor([0, 9, 17, 31])
    ri = synth.to_reference_layout(case)
    sim = port.similarity_search(ri["src_feats"][sel], ri["tar_feat"][sel], ri["src_masks"][sel], ri["tar_mask"][sel])
    eng.set_queries(case.q_feat, case.q_mask16.reshape(-1, 16, 16), case.q_label - 1)
    m = cpu(eng.sim_topk())
    for k in ("id_src", "src_pts", "tar_pts"):
        assert torch.equal(m[k][sel], sim[k]), k

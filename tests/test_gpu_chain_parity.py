"""-m gpu: parity of the WHOLE hot path against the CPU oracle, from crops to poses and on every BASELINE.json
configuration shape.

* crop level (reference `src/models/gigaPose.py:497-604`): the same 224x224 crops go through
  `port.ae_features` + `port.ISTBackbonePort` + `port.retrieval` on the CPU and through `GigaPose.retrieve` (native ViT-L/14,
  native IST trunk, resident bank, fused similarity search, MLP, RANSAC, pose) on the GPU.  For the fp32-faithful
  (`fp32_split`) ViT the template ids and patch correspondences must be EQUAL; rows a7-a9 are compared statistically
  (RANSAC's 14-px inlier test is a knife edge on fp32 noise when the regressor has random weights: identical inlier
  sets and pose <= 1e-3 for most hypotheses, counts within 2 on the rest; given identical inputs the kernels agree
  with the oracle's RANSAC / pose lifting).  The plain-bf16 ViT runs through the same comparison and its flip counts are
  reported (not asserted): the measurement behind keeping 3 tensor passes in the ViT linears.
* feature level, full a4-a9 chain against the oracle on slices of c2 (8 x 162, B=32), c3 (30 x 162, B=64),
  c4 (21 x 162, B=128) and a T=576 bank (the c5 template count).
"""
import os
import sys

import pandas as pd
import pytest
import torch

from gigapose_b200 import synth
from oracle import port

from helpers import INT_KEYS, assert_chain_equal, cpu, engine_from_case, reference_slice, run_engine, write_report

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------------- crop level
def _models_with_oracle_weights():
    """Native modules carrying exactly the oracle's (seeded) weights."""
    import src.megapose.utils.tensor_collection as tc  # noqa: F401
    from gigapose_b200.vit import DinoVisionTransformer
    from src.models.gigaPose import GigaPose
    from src.models.matching import LocalSimilarity
    from src.models.network.ae_net import AENet
    from src.models.network.ist_net import ISTNet, Regressor
    from src.models.network.resnet import ResNet
    ref_vit = port.DinoV2Port(depth=24, seed=7)
    ref_bb = port.ISTBackbonePort()
    ref_reg = port.RegressorPort(seed=12)
    vit = DinoVisionTransformer(depth=24)
    vit.load_state_dict(ref_vit.state_dict())
    ae = AENet("dinov2_vitl14", dinov2_model=vit, descriptor_size=1024, max_batch_size=64)
    backbone = ResNet(dict(n_heads=0, input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512],
                           descriptor_size=256))
    reg = Regressor(descriptor_size=256, hidden_dim=256, use_tanh_act=True, normalize_output=True)
    ist = ISTNet("resnet", backbone, reg, max_batch_size=64)
    ist.backbone.load_state_dict(ref_bb.state_dict())          # ISTNet re-initialises conv / linear layers
    ist.regressor.load_state_dict(ref_reg.state_dict())
    metric = LocalSimilarity(k=5, sim_threshold=0.5, patch_threshold=3)
    model = GigaPose("large", ae, ist, training_loss=None, testing_metric=metric, optim_config=None, log_interval=1000,
                     log_dir=os.path.join(ROOT, "gpurun_out", "test_logs"), max_num_dets_per_forward=None)
    return model.to(DEV).eval(), ref_vit, ref_bb, ref_reg


def _count_flips(out, ref):
    """Differences against the oracle, hypotheses matched by template id (robust to a different inlier-count order)."""
    B, K = ref["id_src"].shape
    missing = pts = pts_flips = inl_sets = 0
    for b in range(B):
        ids = out["id_src"][b].tolist()
        for kk in range(K):
            tid = int(ref["id_src"][b, kk])
            if tid not in ids:
                missing += 1
                continue
            j = ids.index(tid)
            pts += 256
            pts_flips += int(((out["src_pts"][b, j] != ref["src_pts"][b, kk]).any(-1) |
                              (out["tar_pts"][b, j] != ref["tar_pts"][b, kk]).any(-1)).sum())
            inl_sets += int(not torch.equal(out["ransac_src_pts"][b, j], ref["ransac_src_pts"][b, kk]))
    best = lambda d: d["id_src"][torch.arange(B), d["score_src"].argmax(dim=1)]
    return {"templates_missing_from_topk": missing, "hypotheses": B * K, "correspondence_slots_compared": pts,
            "correspondence_flips": pts_flips, "hypotheses_with_other_inlier_set": inl_sets,
            "best_similarity_template_flips": int((best(out) != best(ref)).sum()),
            "score_src_err_max": float((out["score_src"].sort(dim=1).values - ref["score_src"].sort(dim=1).values).abs().max())}


def test_crop_level_chain_matches_oracle():
    sys.path.insert(0, ROOT)
    import bench
    O, T, B = 2, 10, 6
    model, ref_vit, ref_bb, ref_reg = _models_with_oracle_weights()
    templates = bench.SyntheticTemplates(O, T, torch.device("cpu"))
    model.template_datasets = {"synthetic": templates}
    model.test_dataset_name = "synthetic"
    batch, labels, views = bench.make_queries(templates, B, seed=3)

    # ---- CPU oracle from the same crops (gigaPose.py:357-398 onboarding, :497-604 retrieval)
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    rgbs, masks = zip(*[templates.crops(o) for o in range(O)])
    rgb_all = torch.cat(rgbs)                                                  # [O*T,3,224,224]
    src_ae = port.ae_features(ref_vit, rgb_all).reshape(O, T, 1024, 16, 16)
    with torch.no_grad():
        src_ist = ref_bb(rgb_all).reshape(O, T, 256, 16, 16)
        tar_ist = ref_bb(batch.tar_img)
    tar_ae = port.ae_features(ref_vit, batch.tar_img)
    lab = labels - 1
    ref_in = dict(src_feats=src_ae[lab], tar_feat=tar_ae, src_masks=torch.stack(masks)[lab], tar_mask=batch.tar_mask,
                  src_ist=src_ist[lab], tar_ist=tar_ist, tar_label=labels, tar_K=batch.tar_K, tar_M=batch.tar_M,
                  template_K=templates.K.repeat(O, 1, 1), template_Ms=templates.M, template_poses=templates.poses.repeat(O, 1, 1, 1))
    ref = port.retrieval(ref_in, ref_reg)
    n_valid = int((ref["src_pts"][..., 0] != -1).sum())
    assert n_valid > 50 * B, "degenerate workload: almost no valid correspondences"

    # ---- GPU, both ViT precisions through the reference-facing module
    report = {"workload": f"{B} query crops vs {O} objects x {T} templates, ViT-L/14 depth 24, oracle-seeded weights",
              "valid_correspondences": n_valid, "hypotheses": B * 5}
    outs = {}
    for prec in ("fp32_split", "bf16"):
        model.ae_net.precision = prec
        model.engines.clear()
        pred = model.retrieve(batch, "synthetic")
        out = cpu({k: getattr(pred, k) for k in INT_KEYS + ["score_src", "score_pts", "relScale", "relInplane", "M",
                                                            "scores", "pred_poses"]})
        outs[prec] = out
        report[prec] = _count_flips(out, ref)
    model.ae_net.precision = None
    out = outs["fp32_split"]
    # (1) rows a1 -> a4 from crops: template ids and patch correspondences are bit-exact (BASELINE north_star)
    for k in ("id_src", "src_pts", "tar_pts"):
        assert torch.equal(_by_template(out, ref, k), ref[k]), f"crop-level fp32_split: {k}"
    assert torch.allclose(_by_template(out, ref, "score_src"), ref["score_src"], atol=5e-6)
    # (2) rows a6 + a5: floats.  The native trunk is fp32-faithful to ~6e-5 of the feature range (the reference's own GPU
    # path runs these convolutions in TF32, ~1e-3); the regressor outputs inherit that
    valid = ref["src_pts"][..., 0] != -1
    d_scale = float((_by_template(out, ref, "relScale") - ref["relScale"]).abs()[valid].max())
    d_inpl = float((_by_template(out, ref, "relInplane") - ref["relInplane"]).abs()[valid].max())
    r_scale = float(ref["relScale"].abs()[valid].max())
    report["fp32_split"].update(relScale_err_max=d_scale, relScale_abs_max=r_scale, relInplane_err_max=d_inpl)
    assert d_scale < 1e-3 * max(1.0, r_scale) and d_inpl < 2e-3, (d_scale, r_scale, d_inpl)
    # (3) rows a7 - a9 given the SAME inputs: the oracle's RANSAC run on the GPU's own (relScale, relInplane).  With
    # generic float inputs (random-weight regressor) the 14-px inlier test has knife-edge cases even between two fp32
    # implementations of the same formula (the kernel reproduces the reference's operation order without FMA contraction,
    # MKL on the CPU contracts), so: identical inlier sets for all but a few hypotheses, counts within 2, and pose
    # lifting (a9) exact to 1e-3 given the same M
    M_o, failed_o, in_src_o, in_tar_o, in_sc_o = port.ransac(out["src_pts"], out["tar_pts"], out["relScale"], out["relInplane"])
    same_k = (out["ransac_src_pts"] == in_src_o).flatten(2).all(-1)
    cnt_diff = (out["ransac_scores"].sum(-1) - in_sc_o.sum(-1)).abs()
    report["fp32_split"].update(a7_same_inputs_identical_inlier_sets=int(same_k.sum()), a7_same_inputs_max_count_diff=int(cnt_diff.max()))
    # measured on B200 boxes: 28 of 30 identical sets, 26 with the same transform, counts within 1; the bars below leave
    # room for another host CPU deciding a knife-edge case the other way (the oracle side runs on the box's CPU)
    assert int(same_k.sum()) >= int(0.8 * B * 5) and int(cnt_diff.max()) <= 2, report["fp32_split"]
    assert torch.equal(out["idx_failed"][same_k], failed_o[same_k])
    # (the regressor has random weights: |M| reaches 1e4, so the bound is relative to each matrix' largest entry)
    # (two candidates can produce the same inlier set with different transforms, so "same set" does not imply the same
    # winner: the bound must hold for at least 85 % of them)
    dM = (out["M"] - M_o).abs().flatten(2).amax(-1)
    close = (dM <= 2e-5 * M_o.abs().flatten(2).amax(-1) + 2e-3) & same_k
    report["fp32_split"].update(a7_same_inputs_same_transform=int(close.sum()))
    assert int(close.sum()) >= int(0.7 * B * 5), report["fp32_split"]
    poses_o = port.pose_recovery(labels, batch.tar_K, batch.tar_M, out["id_src"], out["M"].clone(), ref_in["template_K"],
                                 ref_in["template_Ms"], ref_in["template_poses"])
    err = (out["pred_poses"] - poses_o).abs()
    err[..., :3, 3] /= poses_o[..., :3, 3].abs().clamp(min=1.0)
    assert float(err.max()) < 1e-3, f"pose lifting error {float(err.max()):.3e}"
    s = out["scores"]
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    # (4) end to end against the oracle, hypotheses matched by template id: RANSAC's 14-px inlier test is a knife edge
    # on fp32 noise, so a small fraction of hypotheses may pick another inlier set; all others agree to 1e-3 on the pose
    same_set = torch.tensor([[torch.equal(_by_template(out, ref, "ransac_src_pts")[b, kk], ref["ransac_src_pts"][b, kk])
                              for kk in range(5)] for b in range(B)])
    perr = (_by_template(out, ref, "pred_poses") - ref["pred_poses"]).abs()
    perr[..., :3, 3] /= ref["pred_poses"][..., :3, 3].abs().clamp(min=1.0)
    perr = perr.flatten(2).amax(-1)
    d_count = (_by_template(out, ref, "ransac_scores").sum(-1) - ref["ransac_scores"].sum(-1)).abs()
    report["fp32_split"].update(hypotheses_with_identical_inlier_set=int(same_set.sum()),
                                max_inlier_count_difference=int(d_count.max()),
                                pose_err_max_identical_inlier_set=float(perr[same_set].max()))
    print("\ncrop-level parity:", report)
    write_report("crop_chain_parity.json", report)
    # (an identical inlier set can still come from another winning candidate, so the pose bound is asked of 85 %)
    good = same_set & (perr < 1e-3)
    report["fp32_split"].update(hypotheses_with_identical_inlier_set_and_pose_1e3=int(good.sum()))
    write_report("crop_chain_parity.json", report)
    assert int(good.sum()) >= int(0.75 * B * 5), report["fp32_split"]          # measured: 27 of 30
    assert int(same_set.sum()) >= int(0.8 * B * 5), report["fp32_split"]       # measured: 28 of 30
    assert int(d_count.max()) <= 2, report["fp32_split"]


def _by_template(out, ref, key):
    """Rows of `out[key]` re-ordered so that hypothesis kk of detection b is the one with template id ref.id_src[b,kk]
    (the two sides sort equal inlier counts / flipped counts differently)."""
    B, K = ref["id_src"].shape
    order = torch.zeros(B, K, dtype=torch.long)
    for b in range(B):
        ids = out["id_src"][b].tolist()
        for kk in range(K):
            tid = int(ref["id_src"][b, kk])
            assert tid in ids, f"detection {b}: template {tid} retrieved by the oracle is missing on the GPU"
            order[b, kk] = ids.index(tid)
    return out[key][torch.arange(B)[:, None], order]


# ------------------------------------------------------------------------------------- feature level, BASELINE shapes
def _chain_on_slice(case, reg, sel, tag, max_batch=None):
    eng = engine_from_case(case, regressor=reg, max_batch=max_batch)
    out = cpu(run_engine(eng, case))
    ref = port.retrieval(reference_slice(case, sel), reg)
    assert_chain_equal(out, ref, sel=sel, tag=tag)
    return out


def test_c2_full_chain_matches_oracle_on_a_slice():
    case = synth.make_feature_case(B=32, O=8, T=162, seed=42)
    _chain_on_slice(case, port.RegressorPort(seed=9), [0, 9, 17, 31], "c2: ")


@pytest.mark.parametrize("name,O,B,sel", [("c3", 30, 64, [1, 40, 63]), ("c4", 21, 128, [0, 77, 127])])
def test_c3_c4_shaped_chain_matches_oracle_on_a_slice(name, O, B, sel):
    """BASELINE.json configs[2], configs[3] on ONE GPU (whole bank resident: 5.1 / 3.6 GB); the bank is generated on
    the device, the oracle sees the selected queries and the full 162-template banks of their objects."""
    case = synth.make_feature_case(B=B, O=O, T=162, seed=50 + O, device=DEV, obj_chunk=2)
    _chain_on_slice(case, port.RegressorPort(seed=10), sel, name + ": ")


def test_t576_chain_matches_oracle():
    """The c5 template count (576 per object): odd number of 32-query chunks per template, 18 top-k tiles."""
    case = synth.make_feature_case(B=6, O=3, T=576, seed=61, device=DEV, obj_chunk=1)
    _chain_on_slice(case, port.RegressorPort(seed=11), [0, 3, 5], "T=576: ")


def test_cuda_graph_with_more_than_one_chunk():
    """B = 2 x max_dets_per_call through the CUDA graph must equal the eager result (every chunk's outputs
    are copied out of the graph's static tensors before the next replay)."""
    sys.path.insert(0, ROOT)
    import bench
    model = bench.build_models(torch.device(DEV))
    model.max_dets_per_call = 4
    templates = bench.SyntheticTemplates(2, 8, torch.device(DEV))
    model.template_datasets = {"synthetic": templates}
    model.test_dataset_name = "synthetic"
    batch, labels, views = bench.make_queries(templates, 8, seed=6)
    model.use_cuda_graph = False
    eager = model.retrieve(batch, "synthetic")
    model.use_cuda_graph = True
    for _ in range(2):
        graphed = model.retrieve(batch, "synthetic")
    for k in ("id_src", "src_pts", "scores", "pred_poses"):
        assert torch.equal(getattr(graphed, k).cpu(), getattr(eager, k).cpu()), k
    # a result stays valid after the next call
    keep = graphed.pred_poses.clone()
    other, _, _ = bench.make_queries(templates, 8, seed=7)
    model.retrieve(other, "synthetic")
    assert torch.equal(graphed.pred_poses, keep)


def test_cuda_graph_batch_size_buckets():
    """Batch sizes 5, 6, 7 share ONE captured graph (padded to 8) and give the eager results."""
    sys.path.insert(0, ROOT)
    import bench
    model = bench.build_models(torch.device(DEV))
    templates = bench.SyntheticTemplates(2, 8, torch.device(DEV))
    model.template_datasets = {"synthetic": templates}
    model.test_dataset_name = "synthetic"
    for B in (5, 6, 7):
        batch, labels, views = bench.make_queries(templates, B, seed=20 + B)
        model.use_cuda_graph = False
        eager = model.retrieve(batch, "synthetic")
        model.use_cuda_graph = True
        graphed = model.retrieve(batch, "synthetic")
        for k in ("id_src", "src_pts", "scores", "pred_poses", "ransac_scores"):
            assert getattr(graphed, k).shape[0] == B
            assert torch.equal(getattr(graphed, k).cpu(), getattr(eager, k).cpu()), (B, k)
    assert len(model._graphs) == 1

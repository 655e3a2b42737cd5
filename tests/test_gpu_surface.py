"""-m gpu tests of the reference-facing Python surface (`src/models/**`): same calls the reference makes, results
against the CPU oracle."""
import pandas as pd
import pytest
import torch

from gigapose_b200 import synth
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case():
    case = synth.make_feature_case(B=4, O=2, T=10, seed=8)
    return case, synth.to_reference_layout(case)


def test_local_similarity_test_signature_and_outputs():
    from src.models.matching import LocalSimilarity
    case, ri = _case()
    metric = LocalSimilarity(k=5, sim_threshold=0.5, patch_threshold=3)
    out = metric.test(src_feats=ri["src_feats"].to(DEV), tar_feat=ri["tar_feat"].to(DEV),
                      src_masks=ri["src_masks"].to(DEV), tar_mask=ri["tar_mask"].to(DEV), max_batch_size=None)
    ref = port.similarity_search(ri["src_feats"], ri["tar_feat"], ri["src_masks"], ri["tar_mask"])
    assert out.id_src.dtype == torch.int64 and out.tar_pts.shape == (4, 5, 256, 2)
    for k in ("id_src", "tar_pts", "src_pts"):
        assert torch.equal(getattr(out, k).cpu(), ref[k]), k
    assert torch.allclose(out.score_src.cpu(), ref["score_src"], atol=2e-6)
    assert torch.allclose(out.score_pts.cpu(), ref["score_pts"], atol=2e-6)


def test_istnet_inference_and_pose_recovery_modules():
    import src.megapose.utils.tensor_collection as tc
    from src.models.network.ist_net import ISTNet, Regressor
    from src.models.network.resnet import ResNet
    from src.models.poses import ObjectPoseRecovery
    case, ri = _case()
    sim = port.similarity_search(ri["src_feats"], ri["tar_feat"], ri["src_masks"], ri["tar_mask"])
    reg_ref = port.RegressorPort(seed=12)
    reg = Regressor(descriptor_size=256, hidden_dim=256, use_tanh_act=True, normalize_output=True)
    reg.load_state_dict(reg_ref.state_dict())
    backbone = ResNet(dict(n_heads=0, input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512],
                           descriptor_size=256))
    ist = ISTNet("resnet", backbone, reg, max_batch_size=64)
    ist.regressor.load_state_dict(reg_ref.state_dict())      # ISTNet re-initialises Linear layers (ist_net.py:33-42)
    ist = ist.to(DEV).eval()
    B, K = sim["id_src"].shape
    bi = torch.arange(B)
    rel_scale = torch.zeros(B, K, 256)
    rel_inpl = torch.zeros(B, K, 256, 2)
    for kk in range(K):
        src_ist = ri["src_ist"][bi, sim["id_src"][:, kk]]
        a, b = ist.inference(src_feat=src_ist.to(DEV), tar_feat=ri["tar_ist"].to(DEV),
                             src_pts=sim["src_pts"][:, kk].to(DEV), tar_pts=sim["tar_pts"][:, kk].to(DEV))
        ra, rb = port.ist_mlp(reg_ref, src_ist, ri["tar_ist"], sim["src_pts"][:, kk], sim["tar_pts"][:, kk])
        assert torch.allclose(a.cpu(), ra, atol=1e-4, rtol=1e-5) and torch.allclose(b.cpu(), rb, atol=1e-4, rtol=1e-5)
        rel_scale[:, kk], rel_inpl[:, kk] = ra, rb
    # ObjectPoseRecovery: forward_ransac on a collection, then forward_recovery
    rec = ObjectPoseRecovery(template_K=ri["template_K"].to(DEV), template_Ms=ri["template_Ms"].to(DEV),
                             template_poses=ri["template_poses"].to(DEV))
    pred = tc.PandasTensorCollection(infos=pd.DataFrame(), src_pts=sim["src_pts"].to(DEV), tar_pts=sim["tar_pts"].to(DEV),
                                     relScale=rel_scale.to(DEV), relInplane=rel_inpl.to(DEV))
    pred = rec.forward_ransac(pred)
    M, failed, in_src, in_tar, in_sc = port.ransac(sim["src_pts"], sim["tar_pts"], rel_scale, rel_inpl)
    assert torch.equal(pred.idx_failed.cpu(), failed)
    assert torch.equal(pred.ransac_scores.cpu(), in_sc) and torch.equal(pred.ransac_src_pts.cpu(), in_src)
    assert torch.allclose(pred.M.cpu(), M, atol=1e-5, rtol=1e-5)
    poses = rec.forward_recovery(tar_label=ri["tar_label"].to(DEV), tar_K=ri["tar_K"].to(DEV), tar_M=ri["tar_M"].to(DEV),
                                 pred_src_views=sim["id_src"].to(DEV), pred_M=pred.M.clone())
    ref = port.pose_recovery(ri["tar_label"], ri["tar_K"], ri["tar_M"], sim["id_src"], M, ri["template_K"],
                             ri["template_Ms"], ri["template_poses"])
    err = (poses.cpu() - ref).abs()
    err[..., :3, 3] /= ref[..., :3, 3].abs().clamp(min=1.0)
    assert float(err.max()) < 1e-3


def test_ransac_module_single_hypothesis():
    import src.megapose.utils.tensor_collection as tc
    from src.models.ransac import RANSAC
    g = torch.Generator().manual_seed(0)
    B, N = 3, 256
    # planted similarity: tar = s * R(theta) * src + t on a subset of patches, exact relScale / relInplane
    src = torch.full((B, N, 2), -1, dtype=torch.long)
    tar = torch.full((B, N, 2), -1, dtype=torch.long)
    rs = torch.full((B, N), -1000.0)
    ri = torch.full((B, N, 2), -1000.0)
    for b in range(B):
        cand = torch.arange(N)[((torch.arange(N) % 16) >= 2) & ((torch.arange(N) % 16) <= 13)]   # shift stays in-grid
        idx = cand[torch.randperm(len(cand), generator=g)[:60]].sort().values
        sx, sy = idx % 16, idx // 16
        shift = torch.tensor([1, -2][b % 2])
        src[b, idx, 0], src[b, idx, 1] = sx, sy
        tar[b, idx, 0], tar[b, idx, 1] = sx + shift, sy
        rs[b, idx] = 1.0
        ri[b, idx, 0], ri[b, idx, 1] = 1.0, 0.0
    batch = tc.PandasTensorCollection(infos=pd.DataFrame(), src_pts=src.to(DEV), tar_pts=tar.to(DEV), relScale=rs.to(DEV),
                                      relInplane=ri.to(DEV))
    Ms, failed, inl = RANSAC(pixel_threshold=14)(batch)
    M_ref, failed_ref, in_src, in_tar, in_sc = port.ransac(src[:, None], tar[:, None], rs[:, None], ri[:, None])
    assert torch.equal(failed.cpu(), failed_ref[:, 0])
    assert torch.allclose(Ms.cpu(), M_ref[:, 0], atol=1e-5)
    assert torch.equal(inl.src_pts.cpu(), in_src[:, 0]) and torch.equal(inl.scores.cpu(), in_sc[:, 0])
    # known answer: pure translation by 14 * shift pixels
    assert torch.allclose(Ms[:, 0, 2].cpu(), torch.tensor([14.0, -28.0, 14.0]))


def test_gigapose_module_end_to_end_small():
    """Crops in, poses out through the reference-facing `GigaPose` surface with synthetic templates."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    model = bench.build_models(torch.device(DEV))
    templates = bench.SyntheticTemplates(2, 8, torch.device(DEV))
    model.template_datasets = {"synthetic": templates}
    model.test_dataset_name = "synthetic"
    batch, labels, views = bench.make_queries(templates, 3, seed=1)
    _, pred = model.eval_retrieval(batch, idx_batch=0, dataset_name="synthetic")
    assert pred.pred_poses.shape == (3, 5, 4, 4) and pred.scores.shape == (3, 5)
    assert torch.isfinite(pred.pred_poses).all()
    # each query is a noisy copy of template `views[b]` of its object: that view must be among the k retrieved
    assert bool((pred.id_src.cpu() == views[:, None]).any(dim=1).all())
    # scores are sorted descending (gigaPose.py:590-595)
    s = pred.scores.cpu()
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    # oracle for the matching stage on the GPU-computed features of the same crops
    eng = model.engines["synthetic"]
    assert eng.launch_count() > 0
    # staged (prefetched on the copy stream) batches give the same answer, eagerly and through the CUDA graph
    for graph in (False, True):
        model.use_cuda_graph = graph
        nxt = model.stage(batch, "synthetic")
        for _ in range(3):
            cur, nxt = nxt, model.stage(batch, "synthetic")
            again = model.retrieve(cur, "synthetic")
            handle = model.fetch_async(again)
            assert torch.equal(again.id_src.cpu(), pred.id_src.cpu())
            poses_host, scores_host = handle.result()
            assert poses_host.is_pinned() and torch.equal(poses_host, pred.pred_poses.cpu())
            assert torch.equal(scores_host, pred.scores.cpu())


def test_ist_backbone_matches_oracle(golden_dir, tol=3e-4):
    """Row a6 against the fp32 CPU oracle's golden features: tcgen05 implicit-GEMM trunk with fp32-faithful split
    products (observed 7e-5 of the feature range after 21 layers).  There is no library / CPU path to fall back to:
    a CPU tensor or another crop size raises."""
    import os
    import numpy as np
    from src.models.network.resnet import ResNet
    ref = port.ISTBackbonePort()
    net = ResNet(dict(n_heads=0, input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512],
                      descriptor_size=256))
    net.load_state_dict(ref.state_dict())
    net = net.to(DEV).eval()
    rgb, _ = synth.make_crops(2, seed=31)
    with torch.no_grad():
        got = net(rgb.to(DEV)).cpu()
    assert getattr(net, "_gp_trunk_engine", None) is not None
    from gigapose_b200._lib import GigaPoseNativeError
    with torch.no_grad():
        with pytest.raises(GigaPoseNativeError):
            net(rgb)                                        # CPU tensor
        with pytest.raises(GigaPoseNativeError):
            net(torch.zeros(1, 3, 256, 256, device=DEV))    # not a 224x224 crop
    g = np.load(os.path.join(golden_dir, "backbones.npz"))
    want = torch.from_numpy(g["ist_feat_sub"])
    scale = want.abs().max().item()
    assert (got[:, ::2] - want).abs().max().item() < tol * scale


def test_test_step_writes_reference_npz_schema(tmp_path):
    """`test_step` with a `test_list` (localisation setting): per-object top-`inst_count` filtering by the best
    hypothesis score and the per-image .npz the reference writes (gigaPose.py:400-449)."""
    import os
    import sys
    import numpy as np
    import pandas as pd
    import src.megapose.utils.tensor_collection as tc
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    model = bench.build_models(torch.device(DEV))
    model.log_dir = str(tmp_path)
    os.makedirs(os.path.join(model.log_dir, "predictions"), exist_ok=True)
    templates = bench.SyntheticTemplates(2, 8, torch.device(DEV))
    model.template_datasets = {"synthetic": templates}
    model.test_dataset_name = "synthetic"
    batch, labels, views = bench.make_queries(templates, 5, seed=4)
    obj_ids = sorted(set(int(l) for l in labels))
    test_list = tc.PandasTensorCollection(infos=pd.DataFrame(dict(obj_id=obj_ids, inst_count=[1] * len(obj_ids),
                                                                  detection_time=[0.25] * len(obj_ids))))
    batch.register_tensor("test_list", test_list)
    assert model.test_step(batch, 7) == 0
    data = np.load(os.path.join(model.log_dir, "predictions", "7.npz"))
    assert set(data.files) == {"scene_id", "im_id", "object_id", "time", "detection_time", "poses", "scores"}
    n = len(obj_ids)                                   # one instance kept per object
    assert data["poses"].shape == (n, 5, 4, 4) and data["scores"].shape == (n, 5)
    assert sorted(data["object_id"].tolist()) == obj_ids
    assert np.allclose(data["detection_time"], 0.25) and (data["time"] > 0).all()
    # the kept detection of each object is the one with the highest top-1 score among that object's detections
    pred = model.retrieve(batch, "synthetic")
    s0 = pred.scores[:, 0].cpu().numpy()
    lab = labels.numpy()
    for row, oid in enumerate(data["object_id"]):
        assert np.isclose(data["scores"][row, 0], s0[lab == oid].max())
    # on_test_epoch_end (gigaPose.py:644-653) turns the per-batch files into the two BOP csv files (row f4)
    from src.utils import inout
    model.run_id = "t0"
    model.on_test_epoch_end()
    stem = os.path.join(model.log_dir, "predictions", f"{model.model_name}-pbrreal-rgb-mmodel_synthetic-test_t0")
    top1 = inout.load_bop_results(stem + ".csv")
    topk = inout.load_bop_results(stem + "MultiHypothesis.csv", additional_name="instance_id")
    assert len(top1) == n and len(topk) == 5 * n
    for row, est in enumerate(top1):
        assert est["obj_id"] == int(data["object_id"][row])
        assert np.allclose(est["R"], data["poses"][row, 0, :3, :3]) and np.allclose(est["t"][:, 0], data["poses"][row, 0, :3, 3])
        assert np.isclose(est["time"], 0.25 + data["time"][row])          # detection time + the one batch that held it


def test_bank_persistence_round_trip(tmp_path):
    """Row f2: the onboarded bank written to disk and read back into a fresh engine gives bit-identical retrieval;
    files for another shape are refused; `GigaPose(bank_cache_dir=...)` skips the template encoders on the second run."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from gigapose_b200.engine import Engine
    from gigapose_b200._lib import GigaPoseNativeError
    dev = torch.device(DEV)
    model = bench.build_models(dev)
    model.bank_cache_dir = str(tmp_path)
    templates = bench.SyntheticTemplates(2, 8, dev)
    model.template_datasets = {"synthetic": templates}
    model.test_dataset_name = "synthetic"
    batch, labels, views = bench.make_queries(templates, 4, seed=2)
    first = model.retrieve(batch, "synthetic")
    files = [f for f in os.listdir(tmp_path) if f.endswith(".gpbank")]
    assert len(files) == 1
    path = os.path.join(tmp_path, files[0])
    eng = model.engines["synthetic"]
    assert os.path.getsize(path) > eng.bank_bytes
    # second model instance: the cache is hit (the template encoders never run) and the results are identical
    model2 = bench.build_models(dev)
    model2.bank_cache_dir = str(tmp_path)
    model2.template_datasets = {"synthetic": templates}
    model2.test_dataset_name = "synthetic"
    calls = {"n": 0}
    orig = model2.ae_net.raw_tokens
    model2.ae_net.raw_tokens = lambda x: (calls.__setitem__("n", calls["n"] + x.shape[0]), orig(x))[1]
    second = model2.retrieve(batch, "synthetic")
    assert calls["n"] == 4                                  # only the 4 query crops went through the ViT
    for name in ("id_src", "pred_poses", "scores"):
        assert torch.equal(getattr(first, name).cpu(), getattr(second, name).cpu()), name
    # a bank of another shape refuses the file
    other = Engine(2, 9, 8, device=DEV)
    with pytest.raises(GigaPoseNativeError, match="does not match"):
        other.load_bank(path)


def test_batched_onboarding_from_raw_renders_equals_the_per_object_path():
    """Row f2: `GigaPose.onboard_templates` (raw RGBA renders + boxes -> `gp_crop_resize_pad` -> both encoders in 64-crop
    chunks that run ACROSS object boundaries -> bank) must produce byte for byte the bank that `set_template_data` builds
    from the crops the reference's `TemplateSet.__getitem__` would hand over (here: the oracle's CropResizePad +
    normalisation, dataloader/template.py:67-73)."""
    import os
    import sys
    import src.megapose.utils.tensor_collection as tc
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from gigapose_b200.preprocess import CLIP_MEAN, CLIP_STD
    dev = torch.device(DEV)
    O, T, H, W = 3, 70, 240, 320                       # 210 crops = 3 x 64 + 18: chunks straddle the objects
    g = torch.Generator().manual_seed(5)
    rgba, boxes, sets = [], [], []
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    K = torch.tensor(synth.LM_K)
    poses = synth.fibonacci_view_poses(T)
    mean, std = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1), torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    for o in range(O):
        tex = torch.nn.functional.interpolate(torch.rand(T, 3, 30, 40, generator=g), size=(H, W), mode="bilinear")
        cx = torch.randint(90, W - 90, (T,), generator=g)
        cy = torch.randint(80, H - 80, (T,), generator=g)
        rad = torch.randint(30, 70, (T,), generator=g)
        alpha = (((xs[None] - cx[:, None, None]) ** 2 + (ys[None] - cy[:, None, None]) ** 2) <= rad[:, None, None] ** 2).float()
        img = torch.cat([tex * alpha[:, None], alpha[:, None]], dim=1)                     # [T,4,H,W] in [0,1]
        box = torch.stack([cx - rad, cy - rad, cx + rad + 1, cy + rad + 1], dim=1)
        ref = port.crop_resize_pad(box, img)                                               # the reference's CropResizePad
        crops = ref["images"].clone()
        crops[:, :3] = (crops[:, :3] - mean) / std                                         # template.py:71-73
        sets.append(tc.PandasTensorCollection(infos=pd.DataFrame(), K=K, rgb=crops[:, :3], mask=crops[:, 3], M=ref["M"],
                                              poses=poses))
        rgba.append(img)
        boxes.append(box)
    model = bench.build_models(dev)
    model.template_datasets = {"per_object": sets}
    model.set_template_data("per_object")
    eng_a = model.engines["per_object"]
    eng_b = model.onboard_templates("raw", rgba, torch.stack(boxes), K, poses.expand(O, T, 4, 4))
    torch.cuda.synchronize()
    assert torch.equal(model.template_datas["raw"].M.cpu(), torch.stack([s.M for s in sets]))
    assert torch.equal(eng_a._bank_view(), eng_b._bank_view()), "banks differ"
    assert model.onboarding_s_per_object > 0

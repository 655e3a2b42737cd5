"""CPU tests of the host-side pieces of the drop-in surface: container semantics, module construction / state-dict
keys the reference checkpoint expects, the synthetic generator."""
import pandas as pd
import pytest
import torch

from gigapose_b200 import synth


def test_tensor_collection_semantics():
    from src.megapose.utils.tensor_collection import PandasTensorCollection, concatenate
    infos = pd.DataFrame(dict(label=["1", "2", "3"], scene_id=[7, 7, 8]))
    c = PandasTensorCollection(infos=infos, a=torch.arange(6).reshape(3, 2), b=torch.ones(3))
    assert len(c) == 3 and c.a.shape == (3, 2)
    sub = c[[2, 0]]
    assert sub.infos.label.tolist() == ["3", "1"] and torch.equal(sub.a, torch.tensor([[4, 5], [0, 1]]))
    c.register_tensor("z", torch.zeros(3))
    c.a = c.a + 1                                    # assignment to a registered name replaces the tensor
    assert c.a[0, 0] == 1 and "z" in c.tensors
    with pytest.raises(AttributeError):
        c.missing
    d = c.clone()
    d.a.zero_()
    assert c.a.sum() != 0
    both = concatenate([c, d])
    assert len(both) == 6 and both.a.shape == (6, 2)
    cat = c.cat_df(d)
    assert cat.b.shape == (6,)


def test_state_dict_keys_match_reference_checkpoint_layout():
    from gigapose_b200.vit import DinoVisionTransformer
    from src.models.network.ae_net import AENet
    from src.models.network.ist_net import ISTNet, Regressor
    from src.models.network.resnet import ResNet
    vit = DinoVisionTransformer(depth=2)
    ae = AENet("dinov2_vitl14", dinov2_model=vit, descriptor_size=1024, max_batch_size=64)
    keys = set(ae.state_dict())
    for k in ("dinov2_model.cls_token", "dinov2_model.pos_embed", "dinov2_model.mask_token",
              "dinov2_model.patch_embed.proj.weight", "dinov2_model.blocks.1.attn.qkv.bias",
              "dinov2_model.blocks.0.ls1.gamma", "dinov2_model.blocks.0.mlp.fc2.weight", "dinov2_model.norm.weight"):
        assert k in keys, k
    backbone = ResNet(dict(n_heads=0, input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512],
                           descriptor_size=256))
    ist = ISTNet("resnet", backbone, Regressor(256, 256, True, True), max_batch_size=64)
    keys = set(ist.state_dict())
    for k in ("backbone.conv1.weight", "backbone.bn1.running_mean", "backbone.layer2.0.downsample.0.weight",
              "backbone.layer4.1.bn2.weight", "backbone.layer4_outconv.weight", "regressor.scale_predictor.4.bias",
              "regressor.inplane_predictor.0.weight"):
        assert k in keys, k
    assert ist.regressor.inplane_predictor[4].weight.shape == (2, 256)


def test_unsupported_configurations_fail_loudly():
    from src.models.matching import LocalSimilarity
    with pytest.raises(NotImplementedError):
        LocalSimilarity(k=5, sim_threshold=0.5, patch_threshold=3, search_direction="src2tar")
    with pytest.raises(NotImplementedError):
        LocalSimilarity(k=5, sim_threshold=0.5, patch_threshold=3, image_size=448)
    m = LocalSimilarity(k=5, sim_threshold=0.5, patch_threshold=3)
    assert (m.k, m.sim_threshold, m.patch_threshold, m.num_patches) == (5, 0.5, 3, 16)


def test_generator_is_deterministic_and_planted():
    a = synth.make_feature_case(B=2, O=2, T=6, seed=4)
    b = synth.make_feature_case(B=2, O=2, T=6, seed=4)
    assert torch.equal(a.bank_feat, b.bank_feat) and torch.equal(a.q_feat, b.q_feat)
    assert torch.allclose(a.bank_feat.norm(dim=-1), torch.ones(2, 6, 256), atol=1e-5)
    # the planted patch correspondence: a template patch shows the base patch its warp index names
    o, tau = int(a.q_label[0]) - 1, int(a.planted["best_template"][int(a.q_label[0]) - 1])
    w = a.planted["warp_index"][o, tau]
    sim = a.q_feat[0] @ a.bank_feat[o, tau].T                    # [t, s]
    s_ok = torch.nonzero(w >= 0)[:, 0]
    assert (sim[w[s_ok], s_ok] > 0.5).float().mean() > 0.95      # planted pairs are well above the 0.5 threshold
    ref = synth.to_reference_layout(a)
    assert ref["src_feats"].shape == (2, 6, 1024, 16, 16) and ref["src_masks"].shape == (2, 6, 224, 224)


def test_ist_trunk_weight_folding_matches_eval_mode_modules():
    """Host logic of the native IST trunk (row a6): BatchNorm folding + [cout,kh,kw,cin] filter layout in the execution
    order `gp_ist_trunk_create` documents, checked on the CPU against the eval-mode torch modules."""
    import torch
    import torch.nn.functional as F
    from gigapose_b200 import ist_trunk
    from src.models.network.resnet import ResNet
    torch.manual_seed(3)
    net = ResNet(dict(n_heads=0, input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512],
                      descriptor_size=256)).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
    assert ist_trunk.supports(net)
    convs = ist_trunk.folded_convs_in_abi_order(net, "cpu")
    assert len(convs) == ist_trunk.NUM_CONVS == 21
    shapes = [tuple(w.shape) for w, _ in convs]
    assert shapes[0] == (128, 7, 7, 3) and shapes[-1] == (256, 1, 1, 512) and convs[-1][1] is None
    assert shapes[5] == (192, 3, 3, 128) and shapes[6] == (192, 1, 1, 128) and shapes[7] == (192, 3, 3, 192)   # conv1, downsample, conv2

    def conv(x, wb, stride, pad):
        w, b = wb
        return F.conv2d(x, w.permute(0, 3, 1, 2), b, stride=stride, padding=pad)

    with torch.no_grad():
        x = torch.randn(1, 3, 64, 64)
        assert torch.allclose(conv(x, convs[0], 2, 3), net.bn1(net.conv1(x)), atol=1e-5)
        t = torch.randn(1, 128, 16, 16)
        blk = net.layer2[0]                                                # the first strided block: entries 5, 6, 7
        y = F.relu(conv(t, convs[5], 2, 1))
        want = F.relu(blk.downsample(t) + blk.bn2(blk.conv2(F.relu(blk.bn1(blk.conv1(t))))))
        got = F.relu(conv(t, convs[6], 2, 0) + conv(y, convs[7], 1, 1))
        assert torch.allclose(got, want, atol=1e-4)
    other = ResNet(dict(n_heads=0, input_dim=3, input_size=256, initial_dim=64, block_dims=[64, 128, 256, 512],
                        descriptor_size=256))
    assert not ist_trunk.supports(other)                                   # other geometries stay on the torch path


def test_labels_outside_the_bank_raise():
    """the reference indexes `ae_features[label - 1]` -- label 0 wraps to the last object, label > O raises;
    here both raise before anything reaches the kernels."""
    import pandas as pd
    import pytest
    from src.models.gigaPose import object_indices
    assert object_indices(pd.DataFrame(dict(label=["1", "3", "2"])), 3).tolist() == [0, 2, 1]
    for bad in (["0", "1"], ["4"], ["-1"]):
        with pytest.raises(IndexError):
            object_indices(pd.DataFrame(dict(label=bad)), 3)


def test_bank_cache_fingerprint_follows_the_weights():
    import torch
    from src.models.gigaPose import weights_fingerprint
    a, b = torch.nn.Linear(8, 8), torch.nn.BatchNorm1d(8)
    f0 = weights_fingerprint(a, b)
    assert f0 == weights_fingerprint(a, b)
    with torch.no_grad():
        a.weight[3, 3] += 1e-3
    assert weights_fingerprint(a, b) != f0
    b.running_var[0] = 2.0
    assert weights_fingerprint(a, b) != f0


def test_detection_windows_and_light_records():
    from gigapose_b200 import multigpu
    assert [multigpu.window(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [multigpu.window(3, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    covered = sorted(i for r in range(8) for i in range(*multigpu.window(128, r, 8)))
    assert covered == list(range(128))
    full, n_full = multigpu.record_layout(4, 5)
    light, n_light = multigpu.record_layout(4, 5, light=True)
    assert set(light) == set(multigpu.LIGHT_FIELDS) and "rel_scale" in full and n_light < n_full
    assert n_light == 4 * 5 * (4 + 4 + 256 * 4 + 256 + 256)            # 1544 B per candidate, no padding needed here
    assert all(off % 16 == 0 for off, _, _ in light.values())


def test_bank_builder_streams_full_chunks_across_objects():
    """Row f2 host logic: the onboarding builder feeds the encoders 64-crop chunks that straddle object boundaries and
    writes every (object, template) slot exactly once, in order (encoders and engine are stand-ins: no GPU needed)."""
    import torch
    from src.models.gigaPose import _BankBuilder

    class Enc:
        def __init__(self):
            self.calls = []

        def raw_tokens(self, rgb):
            self.calls.append(rgb.shape[0])
            return rgb[:, :1, :1, 0].reshape(-1, 1, 1).repeat(1, 257, 1024)      # token value = crop id

        def forward_by_chunk(self, rgb):
            return rgb[:, :1, :1, :1].repeat(1, 256, 16, 16)

    class Model:
        pass

    class Eng:
        def __init__(self):
            self.writes = []

        def bank_write(self, obj, t0, tokens, mask, ist_feat=None, norm_passes=1):
            assert norm_passes == 2 and tokens.shape[1:] == (257, 1024) and ist_feat.shape[1:] == (256, 16, 16)
            ids = tokens[:, 0, 0].long().tolist()
            assert ids == ist_feat[:, 0, 0, 0].long().tolist() == mask[:, 0, 0].long().tolist()
            self.writes.append((obj, t0, ids))

    model, eng = Model(), Eng()
    model.ae_net = Enc()
    model.ist_net = model.ae_net
    b = _BankBuilder(model, eng, chunk=64)
    O, T = 3, 70
    for o in range(O):
        ids = torch.arange(o * T, (o + 1) * T, dtype=torch.float32)
        b.add(o, ids.view(T, 1, 1, 1).expand(T, 3, 4, 4), ids.view(T, 1, 1).expand(T, 4, 4))
    b.flush()
    assert model.ae_net.calls == [64, 64, 64, 18] and b.crops == O * T
    seen = {}
    for obj, t0, ids in eng.writes:
        for j, cid in enumerate(ids):
            assert (obj, t0 + j) not in seen
            seen[(obj, t0 + j)] = cid
    assert seen == {(o, t): o * T + t for o in range(O) for t in range(T)}

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

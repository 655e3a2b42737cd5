"""Rows a6 / f1: the native IST trunk (csrc/ist_trunk.cu: fused resize + stem im2col, 21 implicit-GEMM convolutions on
tcgen05 with 4-D TMA filter taps) against plain torch fp32 modules of the same network, layer by layer, through the C ABI
(`gp_debug_ist_trunk`, `gp_ist_trunk_forward`).  Reference: src/models/network/resnet.py:26-50,318-381."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512], descriptor_size=256, n_heads=0)


def _network(seed=0):
    from src.models.network.resnet import ResNet
    torch.manual_seed(seed)
    net = ResNet(CFG).to(DEV).eval()
    with torch.no_grad():
        for m in net.modules():                      # non-trivial inference statistics so that the folding is exercised
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
    return net


def _torch_activations(net, x):
    """fp32 (TF32 off) activations in the execution order of gp_ist_trunk_create, and the final feature map."""
    acts = []
    t = F.interpolate(x, (256, 256), mode="bilinear", align_corners=True)
    t = F.relu(net.bn1(net.conv1(t))); acts.append(t)
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        for blk in layer:
            y = F.relu(blk.bn1(blk.conv1(t))); acts.append(y)
            sc = t
            if blk.downsample is not None:
                sc = blk.downsample(t); acts.append(sc)
            t = F.relu(sc + blk.bn2(blk.conv2(y))); acts.append(t)
    return acts, net.layer4_outconv(t)


@pytest.fixture()
def fp32_convs():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_every_convolution_matches_torch_fp32(fp32_convs):
    from gigapose_b200.ist_trunk import NativeISTTrunk
    net = _network()
    x = torch.randn(3, 3, 224, 224, device=DEV)
    eng = NativeISTTrunk(net, DEV, max_crops=4)
    with torch.no_grad():
        acts, final = _torch_activations(net, x)
    assert len(acts) == 20
    for i, want in enumerate(acts, 1):               # strides 1 and 2, 3x3 / 1x1 / 7x7, Cout 128 / 192 / 256 / 512
        got = eng.activation_after(x, i).permute(0, 3, 1, 2)
        assert got.shape == want.shape
        err = (got - want).abs().max().item()
        assert err < 3e-4 * max(want.abs().max().item(), 1.0), (i, err)
    out = eng.forward(x)
    assert out.shape == (3, 256, 16, 16)
    assert (out - final).abs().max().item() < 3e-4 * final.abs().max().item()


def test_module_forward_uses_native_trunk_and_chunks(fp32_convs):
    """`ResNet.forward` (the call ISTNet.forward_by_chunk makes, ist_net.py:62-63) with more crops than one engine pass."""
    from gigapose_b200 import ist_trunk
    net = _network(1)
    x = torch.randn(37, 3, 224, 224, device=DEV)
    with torch.no_grad():
        got = net(x)
        assert net._gp_trunk_engine[1].max_crops == 32
        _, want = _torch_activations(net, x)         # fp32 torch convolutions (TF32 off by the fixture)
    assert got.shape == want.shape == (37, 256, 16, 16)
    assert (got - want).abs().max().item() < 3e-4 * want.abs().max().item()
    assert ist_trunk.BACKEND == "native-tcgen05"


def test_bf16_precision_is_one_pass(fp32_convs):
    from gigapose_b200.ist_trunk import NativeISTTrunk
    net = _network(2)
    x = torch.randn(2, 3, 224, 224, device=DEV)
    with torch.no_grad():
        _, final = _torch_activations(net, x)
    out = NativeISTTrunk(net, DEV, max_crops=2, precision="bf16").forward(x)
    rel = (out - final).abs().max().item() / final.abs().max().item()
    assert 3e-4 < rel < 5e-2                         # plain bf16 products: visibly coarser, still the same function

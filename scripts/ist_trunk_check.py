"""Layer-by-layer check of the native IST trunk against torch fp32 convolutions (GPU box only)."""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from src.models.network.resnet import ResNet
from gigapose_b200.ist_trunk import NativeISTTrunk

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = dict(input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512], descriptor_size=256, n_heads=0)
net = ResNet(cfg).to(dev).eval()
with torch.no_grad():
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = torch.randn(n, 3, 224, 224, device=dev)
eng = NativeISTTrunk(net, dev, max_crops=4)

acts = []
with torch.no_grad():
    t = F.interpolate(x, (256, 256), mode="bilinear", align_corners=True)
    t = F.relu(net.bn1(net.conv1(t))); acts.append(t)
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        for blk in layer:
            y = F.relu(blk.bn1(blk.conv1(t))); acts.append(y)
            sc = t
            if blk.downsample is not None:
                sc = blk.downsample(t); acts.append(sc)
            t = F.relu(sc + blk.bn2(blk.conv2(y))); acts.append(t)
    final = net.layer4_outconv(t)
bad = 0
for i, ref in enumerate(acts, 1):
    got = eng.activation_after(x, i).permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    flag = "" if err <= 2e-4 * max(scale, 1.0) else "  <-- MISMATCH"
    bad += bool(flag)
    print(f"conv {i:2d} shape {tuple(ref.shape)} max|ref| {scale:.3f} max err {err:.3e}{flag}", flush=True)
out = eng.forward(x)
err = (out - final).abs().max().item()
print(f"final {tuple(final.shape)} max|ref| {final.abs().max().item():.3f} max err {err:.3e}")
torch.cuda.synchronize()
# timing at the bench batch
eng32 = NativeISTTrunk(net, dev, max_crops=32)
xb = torch.randn(32, 3, 224, 224, device=dev)
for _ in range(3):
    eng32.forward(xb)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    eng32.forward(xb)
ev[1].record(); torch.cuda.synchronize()
print(f"native trunk: {ev[0].elapsed_time(ev[1]) / 10:.3f} ms / 32 crops")
torch.backends.cudnn.allow_tf32 = True
from ist_cudnn_compare import folded_forward  # noqa: E402
with torch.no_grad():
    for _ in range(3):
        folded_forward(net, xb)
    ev[0].record()
    for _ in range(10):
        folded_forward(net, xb)
    ev[1].record(); torch.cuda.synchronize()
print(f"cuDNN (tf32) trunk: {ev[0].elapsed_time(ev[1]) / 10:.3f} ms / 32 crops")
sys.exit(1 if bad or err > 2e-3 else 0)

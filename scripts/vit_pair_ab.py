"""A/B on one box: ViT-L/14 forward (32 crops) with 1-CTA 128x256 tiles vs 2-CTA cluster 256x256 tiles."""
import os, sys
import torch
sys.path.insert(0, ".")
from gigapose_b200.vit import DinoVisionTransformer
from gigapose_b200.vit_engine import NativeViT
dev = torch.device("cuda:0")
vit = DinoVisionTransformer(init_seed=7).to(dev)
os.environ["GIGAPOSE_GEMM_PAIR"] = "0"; e0 = NativeViT(vit, dev, max_crops=32)
os.environ["GIGAPOSE_GEMM_PAIR"] = "1"; e1 = NativeViT(vit, dev, max_crops=32)
x = torch.randn(32, 3, 224, 224, device=dev)
o0 = e0.forward(x); o1 = e1.forward(x); torch.cuda.synchronize()
print("max |pair - single|", (o0 - o1).abs().max().item(), "max |out|", o0.abs().max().item())
def t(e, n=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        e.forward(x)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for e in (e0, e1): t(e, 5)
for rnd in range(4):
    print(f"round {rnd}: single {t(e0):.3f} ms   pair {t(e1):.3f} ms", flush=True)
# same A/B for the IST trunk
from src.models.network.resnet import ResNet
from gigapose_b200.ist_trunk import NativeISTTrunk
cfg = dict(input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512], descriptor_size=256, n_heads=0)
net = ResNet(cfg).to(dev).eval()
os.environ["GIGAPOSE_CONV_PAIR"] = "0"; t0 = NativeISTTrunk(net, dev, max_crops=32)
os.environ["GIGAPOSE_CONV_PAIR"] = "1"; t1 = NativeISTTrunk(net, dev, max_crops=32)
f0 = t0.forward(x); f1 = t1.forward(x); torch.cuda.synchronize()
print("trunk max |pair - single|", (f0 - f1).abs().max().item(), "max |out|", f0.abs().max().item())
for e in (t0, t1): t(e, 5)
for rnd in range(3):
    print(f"trunk round {rnd}: single {t(t0):.3f} ms   pair {t(t1):.3f} ms", flush=True)

"""One forward of the native IST trunk at 32 crops (for an ncu launch list)."""
import sys
import torch
sys.path.insert(0, ".")
from src.models.network.resnet import ResNet
from gigapose_b200.ist_trunk import NativeISTTrunk
dev = torch.device("cuda:0")
cfg = dict(input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512], descriptor_size=256, n_heads=0)
net = ResNet(cfg).to(dev).eval()
eng = NativeISTTrunk(net, dev, max_crops=32)
x = torch.randn(32, 3, 224, 224, device=dev)
eng.forward(x); torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(x); torch.cuda.synchronize()
torch.cuda.profiler.stop()

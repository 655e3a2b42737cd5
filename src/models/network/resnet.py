"""`ResNet` -- the IST backbone (reference `src/models/network/resnet.py:26-50,318-381`; Hydra target
`src.models.network.resnet.ResNet`, configs/model/ist_net/resnet.yaml:6).

Rows a6 / f1 of SURVEY.md §8.  Inference (`eval()`, no grad) runs on the native tcgen05 implicit-GEMM kernels
(`gigapose_b200/ist_trunk.py`, csrc/ist_trunk.cu) and on nothing else: a CPU tensor, another crop size or another
geometry raises instead of falling back to a library path (BASELINE north_star: no multi-backend dispatch, no CPU
fallback).  The cuDNN comparator lives in `scripts/ist_cudnn_compare.py`.  With autograd enabled / in train mode the plain
torch modules run (training is out of scope; the modules exist so that checkpoints load strictly).  Parameter names match the reference state dict (`conv1, bn1, layer{1..4}.{0,1}.{conv1,conv2,bn1,bn2,downsample.{0,1}}, layer4_outconv`).
The optional attention blocks of the reference (n_heads > 0) are dead under the shipped config and not provided.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class BasicBlock(nn.Module):
    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride=stride, padding=1, bias=False)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1:
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes))

    def forward(self, x):
        y = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)))))
        shortcut = x if self.downsample is None else self.downsample(x)
        return F.relu(shortcut + y)


class ResNet(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.get("n_heads", 0) > 0:
            raise NotImplementedError("SpatialTransformer blocks (n_heads > 0) are not part of the shipped config")
        self.input_size = config["input_size"]
        width = config["initial_dim"]
        dims = list(config["block_dims"])
        self.conv1 = nn.Conv2d(config["input_dim"], width, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        stages = []
        for i, d in enumerate(dims):
            stages.append(nn.Sequential(BasicBlock(width, d, stride=1 if i == 0 else 2), BasicBlock(d, d, stride=1)))
            width = d
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        self.layer4_outconv = nn.Conv2d(dims[3], config["descriptor_size"], 1, bias=False)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        inference = not self.training and not torch.is_grad_enabled()
        if inference:
            from gigapose_b200 import _lib, ist_trunk
            if not x.is_cuda:
                raise _lib.GigaPoseNativeError("ResNet inference runs on the sm_100a kernels only: got a CPU tensor (no CPU fallback)")
            if tuple(x.shape[1:]) != (3, 224, 224) or not ist_trunk.supports(self):
                raise _lib.GigaPoseNativeError(
                    f"the IST trunk kernels are specialised for 3x224x224 crops and the shipped geometry "
                    f"(configs/model/ist_net/resnet.yaml); got input {tuple(x.shape)}")
            return ist_trunk.trunk_forward(self, x)          # resize, 21 convolutions, BN, ReLU, shortcuts: all native
        # autograd path (the reference resizes 224 -> 256 with align_corners=True before the trunk, resnet.py:365-368)
        x = F.interpolate(x, (self.input_size, self.input_size), mode="bilinear", align_corners=True)
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.layer4_outconv(x)

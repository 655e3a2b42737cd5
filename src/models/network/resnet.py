"""`ResNet` -- the IST backbone (reference `src/models/network/resnet.py:26-50,318-381`; Hydra target
`src.models.network.resnet.ResNet`, configs/model/ist_net/resnet.yaml:6).

Row a6 / f1 of SURVEY.md §8: this stage is the one allowed to stay on library kernels (cuDNN convolutions through
torch) until its native implicit-GEMM version lands; it is listed as `library` in bench.py.  Parameter names match
the reference state dict (`conv1, bn1, layer{1..4}.{0,1}.{conv1,conv2,bn1,bn2,downsample.{0,1}}, layer4_outconv`).
The optional attention blocks of the reference (n_heads > 0) are dead under the shipped config and not provided.
"""
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
        # the reference resizes 224 -> 256 with align_corners=True before the trunk (resnet.py:365-368)
        x = F.interpolate(x, (self.input_size, self.input_size), mode="bilinear", align_corners=True)
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.layer4_outconv(x)

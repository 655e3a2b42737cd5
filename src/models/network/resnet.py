"""`ResNet` -- the IST backbone (reference `src/models/network/resnet.py:26-50,318-381`; Hydra target
`src.models.network.resnet.ResNet`, configs/model/ist_net/resnet.yaml:6).

Rows a6 / f1 of SURVEY.md §8.  Inference on a CUDA device (`eval()`, no grad, 224x224 crops, shipped geometry) runs
on the native tcgen05 implicit-GEMM kernels (`gigapose_b200/ist_trunk.py`, csrc/ist_trunk.cu); `backend = "cudnn"`
selects the BatchNorm-folded cuDNN path instead (TF32 convolutions; kept for comparison and for other geometries), and
training / autograd uses the plain torch modules.  Parameter names match the reference state dict (`conv1, bn1, layer{1..4}.{0,1}.{conv1,conv2,bn1,bn2,downsample.{0,1}}, layer4_outconv`).
The optional attention blocks of the reference (n_heads > 0) are dead under the shipped config and not provided.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """Inference-time BatchNorm folding: conv(x) * g + h  with g = gamma / sqrt(var + eps), h = beta - mean * g."""
    g = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = (conv.weight * g.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
    return w, (bn.bias - bn.running_mean * g).contiguous()


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
        self.backend = config.get("backend", "native")        # "native" (tcgen05 kernels) | "cudnn"
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
        inference = not self.training and x.is_cuda and not torch.is_grad_enabled()
        if inference and self.backend == "native" and tuple(x.shape[1:]) == (3, 224, 224):
            from gigapose_b200 import ist_trunk
            if ist_trunk.supports(self):
                return ist_trunk.trunk_forward(self, x)      # resize, 21 convolutions, BN, ReLU, shortcuts: all native
        x = F.interpolate(x, (self.input_size, self.input_size), mode="bilinear", align_corners=True)
        if inference:
            return self._forward_folded(x)
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.layer4_outconv(x)

    # ---- inference path: BatchNorm folded into the convolutions, NHWC end to end (no layout-conversion kernels)
    def _folded_params(self):
        key = sum(int(p._version) for p in self.parameters()) + sum(int(b._version) for b in self.buffers())
        cache = getattr(self, "_folded", None)
        if cache is not None and cache[0] == key:
            return cache[1]
        with torch.no_grad():
            params = {"stem": _fold(self.conv1, self.bn1), "blocks": [],
                      "out": self.layer4_outconv.weight.contiguous(memory_format=torch.channels_last)}
            for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
                for blk in layer:
                    ds = _fold(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None
                    params["blocks"].append((_fold(blk.conv1, blk.bn1), _fold(blk.conv2, blk.bn2), ds,
                                             blk.conv1.stride))
        object.__setattr__(self, "_folded", (key, params))
        return params

    def _forward_folded(self, x):
        p = self._folded_params()
        x = x.contiguous(memory_format=torch.channels_last)
        fused = x.is_cuda and hasattr(torch, "cudnn_convolution_relu")
        one = (1, 1)

        def conv_relu(inp, w, b, stride, pad):
            if fused:     # cuDNN's fused conv + bias + ReLU (no separate elementwise kernels)
                return torch.cudnn_convolution_relu(inp, w, b, stride, (pad, pad), one, 1)
            return F.relu_(F.conv2d(inp, w, b, stride=stride, padding=pad))

        w, b = p["stem"]
        x = conv_relu(x, w, b, (2, 2), 3)
        for (w1, b1), (w2, b2), ds, stride in p["blocks"]:
            y = conv_relu(x, w1, b1, tuple(stride), 1)
            if ds is not None:
                x = F.conv2d(x, ds[0], ds[1], stride=stride)
            if fused:     # relu(conv(y) + bias + 1.0 * shortcut)
                x = torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, one, (1, 1), one, 1)
            else:
                x = F.relu_(x + F.conv2d(y, w2, b2, stride=1, padding=1))
        return F.conv2d(x, p["out"]).contiguous()

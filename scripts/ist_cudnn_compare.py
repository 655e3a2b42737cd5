"""Comparator only (NOT a product path): the IST trunk (row a6) on cuDNN through torch, BatchNorm folded, NHWC, fused
conv+bias+ReLU -- what the native tcgen05 trunk is measured against.  `folded_forward(net, x)` takes a
`src.models.network.resnet.ResNet` and 224x224 crops on a CUDA device.

    python scripts/ist_cudnn_compare.py          # accuracy vs fp32 torch + timing of both trunks, 32 crops
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _fold(conv, bn):
    g = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = (conv.weight * g.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
    return w, (bn.bias - bn.running_mean * g).contiguous()


@torch.no_grad()
def folded_forward(net, x):
    x = F.interpolate(x, (net.input_size, net.input_size), mode="bilinear", align_corners=True)
    x = x.contiguous(memory_format=torch.channels_last)
    one = (1, 1)
    w, b = _fold(net.conv1, net.bn1)
    x = torch.cudnn_convolution_relu(x, w, b, (2, 2), (3, 3), one, 1)
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        for blk in layer:
            w1, b1 = _fold(blk.conv1, blk.bn1)
            w2, b2 = _fold(blk.conv2, blk.bn2)
            y = torch.cudnn_convolution_relu(x, w1, b1, tuple(blk.conv1.stride), (1, 1), one, 1)
            if blk.downsample is not None:
                wd, bd = _fold(blk.downsample[0], blk.downsample[1])
                x = F.conv2d(x, wd, bd, stride=blk.conv1.stride)
            x = torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, one, (1, 1), one, 1)
    return F.conv2d(x, net.layer4_outconv.weight.contiguous(memory_format=torch.channels_last)).contiguous()


def main():
    from src.models.network.resnet import ResNet
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = ResNet(dict(n_heads=0, input_dim=3, input_size=256, initial_dim=128, block_dims=[128, 192, 256, 512],
                      descriptor_size=256)).to(dev).eval()
    x = torch.randn(32, 3, 224, 224, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.no_grad():
        native = net(x)
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            lib = folded_forward(net, x)
            for _ in range(3):
                folded_forward(net, x)
            ev[0].record()
            for _ in range(10):
                folded_forward(net, x)
            ev[1].record()
            torch.cuda.synchronize()
            err = (native - lib).abs().max().item() / lib.abs().max().item()
            print(f"cuDNN trunk (tf32={tf32}): {ev[0].elapsed_time(ev[1]) / 10:.3f} ms / 32 crops; native vs cuDNN rel err {err:.2e}")
        for _ in range(3):
            net(x)
        ev[0].record()
        for _ in range(10):
            net(x)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"native tcgen05 trunk: {ev[0].elapsed_time(ev[1]) / 10:.3f} ms / 32 crops")


if __name__ == "__main__":
    main()

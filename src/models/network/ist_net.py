"""`ISTNet` / `Regressor` -- drop-ins for `src/models/network/ist_net.py:11-162` (Hydra targets
configs/model/ist_net/resnet.yaml:1,16).  `inference` (the per-correspondence scale / in-plane MLP of row a5) runs
in the CUDA kernels of `gigapose_b200/csrc/ist_mlp.cu`; the backbone (row a6) is `ResNet`.
"""
import torch
from torch import nn

from src.models._lightning import LightningModule
from src.utils.logging import get_logger

logger = get_logger(__name__)


class Regressor(nn.Module):
    """Two 3-layer MLP heads on cat(query feature, template feature): scale (1) and (cos, sin) in-plane (2)."""

    def __init__(self, descriptor_size, hidden_dim, use_tanh_act, normalize_output):
        super().__init__()
        self.descriptor_size = descriptor_size
        self.normalize_output = normalize_output
        self.use_tanh_act = use_tanh_act
        d, h = descriptor_size * 2, hidden_dim

        def head(out_dim, last):
            return nn.Sequential(nn.Linear(d, 2 * h), nn.ReLU(inplace=True), nn.Linear(2 * h, h), nn.ReLU(inplace=True),
                                 nn.Linear(h, out_dim), *last)

        self.scale_predictor = head(1, [])
        self.inplane_predictor = head(2, [nn.Tanh() if use_tanh_act else nn.Identity()])
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class ISTNet(LightningModule):
    def __init__(self, model_name, backbone, regressor, max_batch_size, patch_size=14, **kwargs):
        super().__init__()
        self.model_name = model_name
        self.patch_size = patch_size
        self.backbone = backbone
        self.regressor = regressor
        self.max_batch_size = max_batch_size
        for module in self.modules():                      # reference ist_net.py:33-42
            if isinstance(module, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(module.weight, mode="fan_in", nonlinearity="relu")
                if module.bias is not None:
                    nn.init.constant_(module.bias, 0)

    def get_toUpdate_parameters(self):
        return list(self.backbone.parameters()) + list(self.regressor.parameters())

    @torch.no_grad()
    def forward_by_chunk(self, processed_rgbs):
        outs = [self.backbone(processed_rgbs[i:i + self.max_batch_size])
                for i in range(0, processed_rgbs.shape[0], self.max_batch_size)]
        return torch.cat(outs, 0) if len(outs) > 1 else outs[0]

    @torch.no_grad()
    def inference(self, src_feat, tar_feat, src_pts, tar_pts):
        """src_feat/tar_feat [B,256,16,16], src_pts/tar_pts [B,N,2] -> scales [B,N], (cos,sin) [B,N,2];
        -1000 where the correspondence is invalid (ist_net.py:97-120)."""
        from gigapose_b200.engine import ist_mlp_explicit
        return ist_mlp_explicit(self.regressor, src_feat, tar_feat, src_pts, tar_pts)

    def inference_by_chunk(self, src_feat, tar_feat, src_pts, tar_pts, max_batch_size):
        return self.inference(src_feat, tar_feat, src_pts, tar_pts)     # the kernels take the batch whole

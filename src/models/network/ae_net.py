"""`AENet` -- drop-in for the reference's `src/models/network/ae_net.py:18-73` (Hydra target
`src.models.network.ae_net.AENet`, configs/model/ae_net/dinov2_l.yaml:1).

Same constructor, same `forward`/`forward_by_chunk` contract: `[b,3,224,224]` -> unit-norm `[b,1024,16,16]` taken
from the PRE-final-norm patch tokens (`x_prenorm[:, 1:]`, ae_net.py:65-69).  The returned tensor is a channels-last
view of a patch-major `[b,256,1024]` buffer: identical values/shape for any caller, and the similarity kernels take
it without a transpose.
"""
import torch

from src.models._lightning import LightningModule
from src.utils.logging import get_logger

logger = get_logger(__name__)
descriptor_sizes = {"dinov2_vits14": 384, "dinov2_vitb14": 768, "dinov2_vitl14": 1024, "dinov2_vitg14": 1536}


class AENet(LightningModule):
    def __init__(self, model_name, dinov2_model, descriptor_size, max_batch_size, patch_size=14, **kwargs):
        super().__init__()
        self.model_name = model_name
        self.dinov2_model = dinov2_model          # state-dict prefix `ae_net.dinov2_model.*`
        self.descriptor_size = descriptor_size
        self.max_batch_size = max_batch_size
        self.patch_size = patch_size
        # "fp32_split" (3-pass bf16 hi/lo products, fp32-faithful) | "bf16" (1 pass); None = GIGAPOSE_VIT_PRECISION / default
        self.precision = kwargs.get("precision", None)

    def get_toUpdate_parameters(self):
        return self.dinov2_model.parameters()

    @torch.no_grad()
    def raw_tokens(self, images: torch.Tensor) -> torch.Tensor:
        """[b,3,224,224] -> `x_prenorm` [b,257,1024] (CLS + patch tokens before the final norm, un-normalised): what the
        resident-bank path hands to `gp_set_queries` / `gp_bank_write` (layout GP_LAYOUT_VIT_TOKENS, norm_passes=2), so
        that the CLS drop and both L2 normalisations (ae_net.py:65-69, matching.py:229) fuse into the plane split."""
        from gigapose_b200.vit_engine import vit_forward_features
        outs = [vit_forward_features(self.dinov2_model, images[i:i + self.max_batch_size], precision=getattr(self, "precision", None))
                for i in range(0, images.shape[0], self.max_batch_size)]
        return torch.cat(outs, 0) if len(outs) > 1 else outs[0]

    @torch.no_grad()
    def patch_tokens(self, images: torch.Tensor) -> torch.Tensor:
        """[b,3,H,W] -> L2-normalised patch-major tokens [b, h*w, C] (ae_net.py:55-69), normalised by the library
        (`gp_normalize_patch_tokens`), not by ATen."""
        from gigapose_b200 import _lib
        tok = self.raw_tokens(images)
        out = torch.empty(tok.shape[0], tok.shape[1] - 1, tok.shape[2], device=tok.device)
        _lib.check(_lib.load().gp_normalize_patch_tokens(tok.shape[0], tok.data_ptr(), out.data_ptr(),
                                                         torch.cuda.current_stream(tok.device).cuda_stream))
        return out

    def forward_by_chunk(self, processed_rgbs, patch_dim=(2, 3)):
        gh = processed_rgbs.shape[patch_dim[0]] // self.patch_size
        gw = processed_rgbs.shape[patch_dim[1]] // self.patch_size
        tok = self.patch_tokens(processed_rgbs)
        return tok.reshape(tok.shape[0], gh, gw, tok.shape[-1]).permute(0, 3, 1, 2)

    def forward(self, images):
        return self.forward_by_chunk(images)

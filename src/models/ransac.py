"""`RANSAC` -- drop-in for `src/models/ransac.py:9-172`: exhaustive one-point similarity-transform search, one CTA
per (detection, hypothesis) in gigapose_b200/csrc/ransac_pose.cu."""
import torch

from src.megapose.utils.tensor_collection import PandasTensorCollection


class RANSAC(torch.nn.Module):
    def __init__(self, pixel_threshold, patch_size=14):
        super().__init__()
        self.patch_size = patch_size
        self.pixel_threshold = pixel_threshold

    @torch.no_grad()
    def forward(self, batch, scores=None, direction="src2tar"):
        from gigapose_b200 import _lib
        from gigapose_b200.engine import ransac_points
        if direction != "src2tar":
            raise NotImplementedError("only direction='src2tar' is used by the inference path (poses.py:146)")
        if scores is not None:
            raise NotImplementedError("per-correspondence weights are always 1 on the inference path (ransac.py:120)")
        src_pts, tar_pts = batch.src_pts.contiguous(), batch.tar_pts.contiguous()
        rel_scale, rel_inplane = batch.relScale.float().contiguous(), batch.relInplane.float()
        if rel_inplane.dim() == 2:        # angle form (ransac.py:81-84): one in-plane angle per correspondence
            rel_inplane = torch.stack([torch.cos(rel_inplane), torch.sin(rel_inplane)], dim=-1)
        rel_inplane = rel_inplane.contiguous()
        B, N = src_pts.shape[:2]
        dev = src_pts.device
        out = dict(M=torch.empty(B, 3, 3, device=dev), idx_failed=torch.empty(B, dtype=torch.uint8, device=dev),
                   ransac_src_pts=torch.empty(B, N, 2, dtype=torch.int64, device=dev),
                   ransac_tar_pts=torch.empty(B, N, 2, dtype=torch.int64, device=dev),
                   ransac_scores=torch.empty(B, N, dtype=torch.int64, device=dev),
                   inlier_count=torch.empty(B, dtype=torch.int32, device=dev))
        ransac_points(_lib.load(), src_pts, tar_pts, rel_scale, rel_inplane, out, self.pixel_threshold, self.patch_size,
                      torch.cuda.current_stream(dev).cuda_stream)
        inliers = PandasTensorCollection(src_pts=out["ransac_src_pts"], tar_pts=out["ransac_tar_pts"],
                                         scores=out["ransac_scores"], infos=batch.infos)
        return out["M"], out["idx_failed"].bool(), inliers

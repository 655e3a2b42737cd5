"""`ObjectPoseRecovery` -- drop-in for `src/models/poses.py:12-163`: RANSAC over the k hypotheses and lifting of
(template id, 2-D similarity, crop matrices, intrinsics) to a 6-D pose, in gigapose_b200/csrc/ransac_pose.cu."""

import torch

from src.models.ransac import RANSAC


class ObjectPoseRecovery(torch.nn.Module):
    def __init__(self, template_K, template_Ms, template_poses, pixel_threshold=14):
        super().__init__()
        self.template_K = template_K.float().contiguous()            # [O,3,3]
        self.template_Ms = template_Ms.float().contiguous()          # [O,T,3,3]
        self.template_poses = template_poses.float().contiguous()    # [O,T,4,4]
        self.ransac = RANSAC(pixel_threshold=pixel_threshold)

    @torch.no_grad()
    def forward_recovery(self, tar_label, tar_K, tar_M, pred_src_views, pred_M):
        """tar_label [B] 1-based object ids (poses.py:111-113), pred_src_views [B,k], pred_M [B,k,3,3] -> [B,k,4,4]."""
        from gigapose_b200 import _lib
        lib = _lib.load()
        dev = pred_M.device
        B, k = pred_src_views.shape
        # this module-level entry point is not on the resident-bank hot path: validating the indices costs one small
        # device->host read (the reference would raise an IndexError from its advanced indexing, poses.py:111-116)
        O, T = self.template_Ms.shape[:2]
        lab = tar_label.to(dev).long() - 1
        views = pred_src_views.to(dev).long()
        bounds = torch.stack([lab.min(), lab.max(), views.min(), views.max()]).tolist() if B * k else [0, 0, 0, 0]
        if bounds[0] < 0 or bounds[1] >= O or bounds[2] < 0 or bounds[3] >= T:
            raise IndexError(f"forward_recovery: object labels must lie in [1, {O}] and template ids in [0, {T}); got labels "
                             f"[{bounds[0] + 1}, {bounds[1] + 1}], template ids [{bounds[2]}, {bounds[3]}]")
        q_obj = lab.to(torch.int32).contiguous()
        tar_K, tar_M = tar_K.to(dev).float().contiguous(), tar_M.to(dev).float().contiguous()
        ids, M = pred_src_views.to(dev).long().contiguous(), pred_M.float().contiguous()
        tK, tM, tP = (t.to(dev) for t in (self.template_K, self.template_Ms, self.template_poses))
        poses = torch.empty(B, k, 4, 4, device=dev)
        _lib.check(lib.gp_pose_recover(B, k, tM.shape[1], q_obj.data_ptr(), tar_K.data_ptr(), tar_M.data_ptr(),
                                       ids.data_ptr(), M.data_ptr(), tK.data_ptr(), tM.data_ptr(), tP.data_ptr(),
                                       poses.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        self._keep = (q_obj, tar_K, tar_M, ids, M, tK, tM, tP)       # alive until the stream has consumed them
        return poses

    @torch.no_grad()
    def forward_ransac(self, predictions):
        """All k hypotheses in one launch (the reference loops over k and over B in python, poses.py:134-147)."""
        from gigapose_b200 import _lib
        from gigapose_b200.engine import ransac_points
        src_pts, tar_pts = predictions.src_pts.contiguous(), predictions.tar_pts.contiguous()
        B, k, N = src_pts.shape[:3]
        dev = src_pts.device
        out = dict(M=torch.empty(B, k, 3, 3, device=dev), idx_failed=torch.empty(B, k, dtype=torch.uint8, device=dev),
                   ransac_src_pts=torch.empty(B, k, N, 2, dtype=torch.int64, device=dev),
                   ransac_tar_pts=torch.empty(B, k, N, 2, dtype=torch.int64, device=dev),
                   ransac_scores=torch.empty(B, k, N, dtype=torch.int64, device=dev),
                   inlier_count=torch.empty(B, k, dtype=torch.int32, device=dev))
        ransac_points(_lib.load(), src_pts, tar_pts, predictions.relScale.float().contiguous(),
                      predictions.relInplane.float().contiguous(), out, self.ransac.pixel_threshold,
                      self.ransac.patch_size, torch.cuda.current_stream(dev).cuda_stream)
        predictions.register_tensor("idx_failed", out["idx_failed"].bool())
        predictions.register_tensor("M", out["M"])
        predictions.register_tensor("ransac_scores", out["ransac_scores"])
        predictions.register_tensor("ransac_src_pts", out["ransac_src_pts"])
        predictions.register_tensor("ransac_tar_pts", out["ransac_tar_pts"])
        return predictions

"""`LocalSimilarity` -- drop-in for `src/models/matching.py:9-316` (Hydra target
`src.models.matching.LocalSimilarity`, configs/model/large.yaml:35-39).

`test()` keeps the reference signature and returns the same `PandasTensorCollection`
(id_src i64 [B,k], score_src f32 [B,k], score_pts f32 [B,k,256], tar_pts / src_pts i64 [B,k,256,2]); the work is
one fused TMA + tcgen05 kernel (gigapose_b200/csrc/sim_search.cu) plus a top-k kernel.  `search()` is the
resident-bank form used by `GigaPose.eval_retrieval`: it skips the reference's 170 MB/detection gather
(gigaPose.py:520-521) by indexing the bank in place.
"""
import pandas as pd
import torch

import src.megapose.utils.tensor_collection as tc


class LocalSimilarity(torch.nn.Module):
    def __init__(self, k, sim_threshold, patch_threshold, search_direction="tar2src", image_size=224, patch_size=14,
                 max_batch_size=32):
        super().__init__()
        if search_direction != "tar2src":
            raise NotImplementedError("only search_direction='tar2src' (the shipped configuration) is implemented")
        if image_size // patch_size != 16:
            raise NotImplementedError("kernels are specialised for a 16x16 patch grid (224 / 14)")
        if not patch_threshold > 0:
            raise NotImplementedError("patch_threshold must be > 0 (cycle-consistency check is always on)")
        self.max_batch_size = max_batch_size          # kept for API compatibility; batches are not chunked
        self.k = k
        self.sim_threshold = sim_threshold
        self.patch_threshold = patch_threshold
        self.search_direction = search_direction
        self.num_patches = image_size // patch_size
        self.precision = "fp32_split"

    @staticmethod
    def _collection(m):
        return tc.PandasTensorCollection(infos=pd.DataFrame(), id_src=m["id_src"], score_src=m["score_src"],
                                         score_pts=m["score_pts"], tar_pts=m["tar_pts"], src_pts=m["src_pts"])

    @torch.no_grad()
    def test(self, src_feats, tar_feat, src_masks, tar_mask, max_batch_size=None):
        """src_feats [B,N,C,16,16] templates, tar_feat [B,C,16,16] queries, masks at image resolution."""
        from gigapose_b200.engine import similarity_search_explicit
        m = similarity_search_explicit(src_feats, tar_feat, src_masks, tar_mask, self.k, self.sim_threshold,
                                       self.patch_threshold, precision=self.precision)
        return self._collection(m)

    @torch.no_grad()
    def search(self, engine):
        """Queries already staged on `engine` (Engine.set_queries) against its resident template bank."""
        return self._collection(engine.sim_topk())

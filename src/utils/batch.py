"""Boundary container of the reference (`src/utils/batch.py:6-43`): a tensor/array/list with a chunk view.

Kept for API compatibility (the reference modules exchange `BatchedData`); the B200 path never chunks through
it -- batches go to the kernels whole.  `gather` (reference :46-73) is provided for callers that still use it.
"""
from __future__ import annotations

import math

import numpy as np
import torch


class BatchedData:
    def __init__(self, batch_size, data=None, **kwargs) -> None:
        self.batch_size = batch_size
        self.data = data if data is not None else []

    def _require_chunking(self):
        if self.batch_size is None:
            raise AssertionError("batch_size is not defined")

    def __len__(self):
        self._require_chunking()
        if isinstance(self.data, (np.ndarray, torch.Tensor)):
            return int(math.ceil(self.data.shape[0] / self.batch_size))
        raise NotImplementedError

    def __getitem__(self, idx):
        self._require_chunking()
        lo = idx * self.batch_size
        return self.data[lo: lo + self.batch_size]

    def cat(self, data, dim=0):
        self.data = data if len(self.data) == 0 else torch.cat([self.data, data], dim=dim)

    def append(self, data):
        self.data.append(data)

    def stack(self, dim=0):
        self.data = torch.stack(self.data, dim=dim)


def gather(features: torch.Tensor, index_patches: torch.Tensor) -> torch.Tensor:
    """features [B,C,H,W], index_patches [B,N,2] as (x,y) with -1 = invalid -> [n_valid, C] (row-major order)."""
    B, C, H, W = features.shape
    valid = (index_patches[..., 0] != -1) & (index_patches[..., 1] != -1)
    b, n = torch.nonzero(valid, as_tuple=True)
    x, y = index_patches[b, n, 0], index_patches[b, n, 1]
    return features[b, :, y, x]


# names this file does not provide resolve from a reference checkout's copy of the same file (see src/__init__.py)
import src as _src  # noqa: E402

__getattr__ = _src.fallback_getattr(__name__)

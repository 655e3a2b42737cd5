"""Boundary I/O structs of the hot path (reference `src/megapose/utils/tensor_collection.py:45,128`):
a named set of tensors sharing their first dimension, optionally with a pandas `infos` frame.

Semantics kept from the reference: attribute access to tensors, `register_tensor`, `__getitem__` on every
tensor (+ `infos.iloc`), `.to()/.cuda()/.cpu()`, `cat_df`, `clone`, `len() == len(infos)`, pickling.
"""
from __future__ import annotations

import pandas as pd
import torch


class TensorCollection:
    def __init__(self, **tensors):
        object.__setattr__(self, "_tensors", {})
        for name, value in tensors.items():
            self.register_tensor(name, value)

    # ---- registry
    def register_tensor(self, name, tensor):
        self._tensors[name] = tensor

    def delete_tensor(self, name):
        del self._tensors[name]

    @property
    def tensors(self):
        return self._tensors

    @property
    def device(self):
        return next(iter(self._tensors.values())).device

    # ---- attribute protocol: tensors first, then plain attributes
    def __getattr__(self, name):
        tensors = object.__getattribute__(self, "_tensors")
        if name in tensors:
            return tensors[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._tensors:
            self._tensors[name] = value
        else:
            object.__setattr__(self, name, value)

    def __getitem__(self, ids):
        return TensorCollection(**{k: v[ids] for k, v in self._tensors.items()})

    def __repr__(self):
        rows = "".join(f"    {k}: {tuple(v.shape)} {v.dtype} {v.device},\n" for k, v in self._tensors.items())
        return f"{type(self).__name__}(\n{rows})"

    # ---- device / dtype moves (in place, like the reference)
    def to(self, target):
        for k in list(self._tensors):
            self._tensors[k] = self._tensors[k].to(target)
        return self

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    def float(self):
        return self.to(torch.float)

    def double(self):
        return self.to(torch.double)

    def half(self):
        return self.to(torch.half)

    def clone(self):
        return TensorCollection(**{k: v.clone() for k, v in self._tensors.items()})

    def __getstate__(self):
        return {"tensors": self._tensors}

    def __setstate__(self, state):
        self.__init__(**state["tensors"])


class PandasTensorCollection(TensorCollection):
    def __init__(self, infos, **tensors):
        super().__init__(**tensors)
        self.infos = infos.reset_index(drop=True)
        self.meta = dict()

    def __len__(self):
        return len(self.infos)

    def __getitem__(self, ids):
        picked = super().__getitem__(ids).tensors
        return PandasTensorCollection(self.infos.iloc[ids].reset_index(drop=True), **picked)

    def clone(self):
        return PandasTensorCollection(self.infos.copy(), **super().clone().tensors)

    def cat_df(self, other):
        for k in list(self._tensors):
            self._tensors[k] = torch.cat([self._tensors[k], other._tensors[k]], dim=0)
        return PandasTensorCollection(infos=self.infos, **self._tensors)

    def cat_df_and_infos(self, other):
        merged = self.cat_df(other)
        infos = pd.concat([self.infos, other.infos], ignore_index=True)
        return PandasTensorCollection(infos=infos, **merged.tensors)

    def merge_df(self, df, *args, **kwargs):
        infos = self.infos.merge(df, how="left", *args, **kwargs)
        assert len(infos) == len(self.infos)
        return PandasTensorCollection(infos=infos, **self._tensors)

    def __repr__(self):
        rows = "".join(f"    {k}: {tuple(v.shape)} {v.dtype} {v.device},\n" for k, v in self._tensors.items())
        return f"{type(self).__name__}(\n{rows}{'-' * 40}\n    infos:\n{self.infos!r}\n)"

    def __getstate__(self):
        state = super().__getstate__()
        state.update(infos=self.infos, meta=self.meta)
        return state

    def __setstate__(self, state):
        self.__init__(state["infos"], **state["tensors"])
        self.meta = state["meta"]


def concatenate(datas):
    datas = [d for d in datas if len(d) > 0]
    if not datas:
        return PandasTensorCollection(infos=pd.DataFrame())
    infos = pd.concat([d.infos for d in datas], axis=0, sort=False).reset_index(drop=True)
    tensors = {k: torch.cat([getattr(d, k) for d in datas], dim=0) for k in datas[0].tensors}
    return PandasTensorCollection(infos=infos, **tensors)


# names this file does not provide resolve from a reference checkout's copy of the same file (see src/__init__.py)
import src as _src  # noqa: E402

__getattr__ = _src.fallback_getattr(__name__)

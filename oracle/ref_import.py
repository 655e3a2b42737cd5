"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Imports the *unmodified* GigaPose reference hot-path modules from /root/reference on CPU.
Only usable in the build container (the reference tree does not travel to the GPU box);
it is what pins `oracle/port.py` and what generates `tests/golden/*.npz`
(see `oracle/make_golden.py`).

Tiny stub modules (pytorch_lightning, omegaconf; bop_toolkit_lib for the writers) are injected because those packages are
not installed here; the reference files themselves are imported as they lie
(SURVEY.md Appendix A).  The repo has its own top-level `src` package (the drop-in surface),
so the reference's `src.*` modules are imported under a temporary sys.modules swap and
returned in a namespace; the repo's `src` is restored afterwards.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("GIGAPOSE_REFERENCE_ROOT", "/root/reference")

_CACHE = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "models"))


def _stub_modules():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        @property
        def device(self):
            p = next(self.parameters(), None)
            return p.device if p is not None else torch.device("cpu")

    pl.LightningModule = LightningModule
    lg = types.ModuleType("pytorch_lightning.loggers")
    lg.WandbLogger = type("WandbLogger", (), {})
    lg.TensorBoardLogger = type("TensorBoardLogger", (), {})
    pl.loggers = lg
    oc = types.ModuleType("omegaconf")
    oc.OmegaConf = type("OmegaConf", (), {})
    oc.DictConfig = dict
    dc = types.ModuleType("omegaconf.dictconfig")
    dc.DictConfig = dict
    oc.dictconfig = dc
    stubs = {"pytorch_lightning": pl, "pytorch_lightning.loggers": lg,
             "omegaconf": oc, "omegaconf.dictconfig": dc}
    try:
        import wandb  # noqa: F401
    except Exception:  # pragma: no cover
        stubs["wandb"] = types.ModuleType("wandb")
    return stubs


class _ReferenceImports:
    """Context manager: sys.modules / sys.path swapped so that `src.*` and `megapose.*` resolve to the reference tree,
    with the stub modules injected; everything is restored on exit."""

    def __init__(self, extra_stubs=None):
        self.extra_stubs = extra_stubs or {}

    @staticmethod
    def is_ref_name(name):
        return name == "src" or name.startswith("src.") or name == "megapose" or name.startswith("megapose.")

    def __enter__(self):
        if not available():
            raise RuntimeError(f"reference tree not found at {REF_ROOT}")
        self.nthreads = torch.get_num_threads()
        self.env_backup = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
        self.saved = {k: v for k, v in sys.modules.items() if self.is_ref_name(k)}
        for k in self.saved:
            del sys.modules[k]
        stubs = dict(_stub_modules(), **self.extra_stubs)
        self.saved_stub = {k: sys.modules.get(k) for k in stubs}
        sys.modules.update(stubs)
        self.old_path = list(sys.path)
        sys.path[:0] = [REF_ROOT, os.path.join(REF_ROOT, "src")]
        # The reference's `src` is a namespace package (no __init__.py) while this repository's `src` is a regular one,
        # and regular packages win regardless of sys.path order: pin `src` to the reference directory explicitly.
        ref_src = types.ModuleType("src")
        ref_src.__path__ = [os.path.join(REF_ROOT, "src")]
        sys.modules["src"] = ref_src
        return self

    def import_reference(self, name):
        m = importlib.import_module(name)
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF_ROOT)), m.__file__
        return m

    def __exit__(self, *exc):
        sys.path[:] = self.old_path
        for k in [k for k in sys.modules if self.is_ref_name(k)]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        for k, v in self.saved_stub.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        # src/megapose/__init__.py:38-39 forces single-threaded BLAS via the environment
        for k, v in self.env_backup.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        torch.set_num_threads(self.nthreads)
        return False


def load():
    """Returns a namespace with the reference classes (LocalSimilarity, RANSAC, ObjectPoseRecovery,
    ISTNet, Regressor, ResNet, AENet, PandasTensorCollection, gather, BatchedData, lib3d)."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    with _ReferenceImports() as ctx:
        ns = types.SimpleNamespace()
        ns.LocalSimilarity = ctx.import_reference("src.models.matching").LocalSimilarity
        ns.RANSAC = ctx.import_reference("src.models.ransac").RANSAC
        ns.ObjectPoseRecovery = ctx.import_reference("src.models.poses").ObjectPoseRecovery
        m = ctx.import_reference("src.models.network.ist_net")
        ns.ISTNet, ns.Regressor = m.ISTNet, m.Regressor
        ns.ResNet = ctx.import_reference("src.models.network.resnet").ResNet
        ns.AENet = ctx.import_reference("src.models.network.ae_net").AENet
        m = ctx.import_reference("src.utils.batch")
        ns.BatchedData, ns.gather = m.BatchedData, m.gather
        ns.lib3d = ctx.import_reference("src.lib3d.torch")
        m = ctx.import_reference("src.megapose.utils.tensor_collection")
        ns.PandasTensorCollection = m.PandasTensorCollection
        ns.tc = m
    _CACHE = ns
    return ns


def load_inout():
    """The reference's result writers (`src/utils/inout.py`, row f4), unmodified.  `bop_toolkit_lib` is not installed
    here and is only imported, never called, on the csv-export path: a stub module stands in for it."""
    bop = types.ModuleType("bop_toolkit_lib")
    bop.inout = types.ModuleType("bop_toolkit_lib.inout")
    with _ReferenceImports({"bop_toolkit_lib": bop, "bop_toolkit_lib.inout": bop.inout}) as ctx:
        return ctx.import_reference("src.utils.inout")

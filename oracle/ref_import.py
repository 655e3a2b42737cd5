"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Imports the *unmodified* GigaPose reference hot-path modules from /root/reference on CPU.
Only usable in the build container (the reference tree does not travel to the GPU box);
it is what pins `oracle/port.py` and what generates `tests/golden/*.npz`
(see `oracle/make_golden.py`).

Two tiny stub modules (pytorch_lightning, omegaconf) are injected because those packages are
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


def load():
    """Returns a namespace with the reference classes (LocalSimilarity, RANSAC, ObjectPoseRecovery,
    ISTNet, Regressor, ResNet, AENet, PandasTensorCollection, gather, BatchedData, lib3d)."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    nthreads = torch.get_num_threads()
    env_backup = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}

    def is_ref_name(name):
        return name == "src" or name.startswith("src.") or name == "megapose" or name.startswith("megapose.")

    saved = {k: v for k, v in sys.modules.items() if is_ref_name(k)}
    for k in saved:
        del sys.modules[k]
    stubs = _stub_modules()
    saved_stub = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    old_path = list(sys.path)
    sys.path[:0] = [REF_ROOT, os.path.join(REF_ROOT, "src")]
    # The reference's `src` is a namespace package (no __init__.py) while this repository's `src` is a regular one,
    # and regular packages win regardless of sys.path order: pin `src` to the reference directory explicitly.
    ref_src = types.ModuleType("src")
    ref_src.__path__ = [os.path.join(REF_ROOT, "src")]
    sys.modules["src"] = ref_src
    try:
        ns = types.SimpleNamespace()
        m = importlib.import_module("src.models.matching")
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF_ROOT)), m.__file__
        ns.LocalSimilarity = m.LocalSimilarity
        m = importlib.import_module("src.models.ransac")
        ns.RANSAC = m.RANSAC
        m = importlib.import_module("src.models.poses")
        ns.ObjectPoseRecovery = m.ObjectPoseRecovery
        m = importlib.import_module("src.models.network.ist_net")
        ns.ISTNet, ns.Regressor = m.ISTNet, m.Regressor
        m = importlib.import_module("src.models.network.resnet")
        ns.ResNet = m.ResNet
        m = importlib.import_module("src.models.network.ae_net")
        ns.AENet = m.AENet
        m = importlib.import_module("src.utils.batch")
        ns.BatchedData, ns.gather = m.BatchedData, m.gather
        ns.lib3d = importlib.import_module("src.lib3d.torch")
        m = importlib.import_module("src.megapose.utils.tensor_collection")
        ns.PandasTensorCollection = m.PandasTensorCollection
        ns.tc = m
    finally:
        sys.path[:] = old_path
        for k in [k for k in sys.modules if is_ref_name(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
        for k, v in saved_stub.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        # src/megapose/__init__.py:38-39 forces single-threaded BLAS via the environment
        for k, v in env_backup.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        torch.set_num_threads(nthreads)
    _CACHE = ns
    return ns

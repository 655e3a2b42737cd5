"""CPU-side checks of the drop-in boundary: the shared library builds/loads without a GPU and exports exactly the
symbols include/gigapose_b200.h declares; configuration errors are reported through the status/last-error channel."""
import ctypes as C
import os
import re

import pytest

from gigapose_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "gigapose_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    declared = _declared_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gigapose_b200.h but not exported"


def test_config_validation_needs_no_gpu(lib):
    cfg = _lib.GpConfig(abi_version=_lib.GP_ABI_VERSION, device=0, num_objects=8, num_templates=162,
                        num_templates_global=162, template_id_stride=1, template_id_offset=0, max_batch=32, top_k=5,
                        sim_threshold=0.5, patch_threshold=3, pixel_threshold=14, patch_size=14, precision=0)
    bank, ws = C.c_size_t(), C.c_size_t()
    assert lib.gp_query_sizes(C.byref(cfg), C.byref(bank), C.byref(ws)) == 0
    # hi + lo bf16 planes == the fp32 bank of BASELINE.md (1.36 GB for 8 x 162) plus masks / IST features / poses
    assert bank.value >= 8 * 162 * 256 * 1024 * 4
    assert bank.value < 1.3 * (8 * 162 * 256 * (1024 * 4 + 256 * 4 + 4))
    cfg.top_k = 0
    assert lib.gp_query_sizes(C.byref(cfg), C.byref(bank), C.byref(ws)) == -1
    assert b"top_k" in lib.gp_last_error()
    cfg.top_k = 5
    cfg.num_templates_global = 3
    assert lib.gp_query_sizes(C.byref(cfg), C.byref(bank), C.byref(ws)) == -1
    cfg.num_templates_global = 162
    cfg.abi_version = 99
    assert lib.gp_query_sizes(C.byref(cfg), C.byref(bank), C.byref(ws)) == -1
    assert b"ABI" in lib.gp_last_error()


def test_engine_refuses_cpu():
    from gigapose_b200.engine import Engine
    with pytest.raises(_lib.GigaPoseNativeError):
        Engine(1, 8, 1, device="cpu")


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure; a product path routed through it would void every parity claim."""
    offenders = []
    for pkg in ("gigapose_b200", "src"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(".py"):
                    txt = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_abi_v2_config_and_argument_checks_need_no_gpu(lib):
    """ABI 2: the replicated IST bank is sized by `ist_bank_global`; the multi-GPU and helper entry points reject bad
    arguments through the status channel without touching a device."""
    def sizes(**kw):
        cfg = _lib.GpConfig(abi_version=_lib.GP_ABI_VERSION, device=0, num_objects=21, num_templates=21,
                            num_templates_global=162, template_id_stride=8, template_id_offset=3, max_batch=128, top_k=5,
                            sim_threshold=0.5, patch_threshold=3, pixel_threshold=14, patch_size=14, precision=0, **kw)
        bank, ws = C.c_size_t(), C.c_size_t()
        rc = lib.gp_query_sizes(C.byref(cfg), C.byref(bank), C.byref(ws))
        return rc, bank.value, ws.value
    rc0, bank_local, ws0 = sizes(ist_bank_global=0)
    rc1, bank_global, ws1 = sizes(ist_bank_global=1)
    assert rc0 == 0 and rc1 == 0 and ws0 == ws1
    extra = 21 * (162 - 21) * 256 * 256 * 4                       # the other shards' IST features, f32 patch-major
    assert extra <= bank_global - bank_local < extra + 4096
    assert sizes(ist_bank_global=2)[0] == -1 and b"ist_bank_global" in lib.gp_last_error()
    assert lib.gp_comm_init(None, None, 0, 1) == -1
    assert lib.gp_allgather(None, None, None, 0, None) == -1
    assert lib.gp_normalize_patch_tokens(0, None, None, None) == -1
    assert lib.gp_bank_write_ist(None, 0, 0, 1, None, 0, None) == -1

"""The C-ABI library loads and exports every symbol include/phylocsf_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from phylocsfpp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "phylocsf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcsf_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(capi.EXPORTS) == syms


def test_abi_version_and_error_string():
    lib = capi.load()
    assert lib.pcsf_abi_version() == 1
    assert isinstance(lib.pcsf_last_error(), bytes)


def test_invalid_arguments_fail_without_a_gpu():
    lib = capi.load()
    h = ctypes.c_void_p()
    assert lib.pcsf_model_create(None, 0, ctypes.byref(h)) == capi.PCSF_ERR_INVALID
    assert b"null" in lib.pcsf_last_error()
    assert lib.pcsf_set_timing(None, 1) == capi.PCSF_ERR_INVALID


def test_no_cpu_fallback():
    """Without a CUDA device model creation must fail loudly (status PCSF_ERR_CUDA), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    from phylocsfpp_b200.models import load_model
    try:
        capi.DeviceModel(load_model("7yeast"), 0)
    except capi.PcsfError as e:
        assert e.status == capi.PCSF_ERR_CUDA
    else:
        raise AssertionError("model creation succeeded without a GPU")

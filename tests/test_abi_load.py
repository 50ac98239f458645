"""CPU-only: the C-ABI library loads and exports every symbol include/fruid_b200.h declares,
the Python binding covers the same set, and -- without a GPU -- calls fail loudly instead of
falling back to a CPU path."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fruid_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fruid_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("fruid_fluid_create", "fruid_fluid_step", "fruid_fluid_upload", "fruid_fluid_download",
                 "fruid_density_create", "fruid_density_step_dev", "fruid_density_step_host",
                 "fruid_op_lin_solve", "fruid_op_advect", "fruid_op_project", "fruid_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from fruid_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(raw, s)]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fruid_b200 as fb
    with pytest.raises(fb.FruidError) as e:
        fb.FluidSim(16, 16)
    assert e.value.code == fb._lib.ERR_CUDA


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "fruid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "fruid_oracle" not in src, f

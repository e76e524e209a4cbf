"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/mfgp.h declares.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from multi_fidelity_gpflow_b200 import build

    return build.build()


def test_header_symbols_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "mfgp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mfgp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mfgp.h but not exported"


def test_binding_covers_header(lib_path):
    from multi_fidelity_gpflow_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "mfgp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mfgp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert _lib._lib.mfgp_version() >= 100


def test_no_cpu_fallback(lib_path):
    """Without a CUDA device a handle cannot be created -- the product never computes on the CPU."""
    import torch

    from multi_fidelity_gpflow_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MFGPError):
        _lib.Handle(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multi_fidelity_gpflow_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), f"{f} references the oracle"

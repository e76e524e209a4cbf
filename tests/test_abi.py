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


class _OnlyDLPack:
    """An exporter that offers nothing but the DLPack protocol (stands in for a TF / JAX tensor)."""

    def __init__(self, base):
        self._base = base
        self.shape = tuple(base.shape)

    def __dlpack__(self, stream=None):
        return self._base.__dlpack__()

    def __dlpack_device__(self):
        return self._base.__dlpack_device__()


def test_dlpack_consumer_host_exporters(lib_path):
    """SURVEY 8(b) data exchange: raw pointers obtained from DLPack capsules (kDLFloat64, C-contiguous, strides checked)."""
    import numpy as np
    import torch

    from multi_fidelity_gpflow_b200 import _lib

    a = np.arange(24, dtype=np.float64).reshape(4, 6)
    p = _lib._ptr(_OnlyDLPack(a))
    assert p.value == a.ctypes.data and _lib.as_f64(_OnlyDLPack(a)).shape == (4, 6)
    addr, shape, kind = _lib.from_dlpack_capsule(a.__dlpack__())
    assert (addr, shape, kind) == (a.ctypes.data, (4, 6), "cpu")
    t = torch.arange(10, dtype=torch.float64)[2:]  # non-zero storage offset: data pointer + byte_offset
    assert _lib._ptr(_OnlyDLPack(t)).value == t.data_ptr()
    assert _lib._ptr(t.__dlpack__()).value == t.data_ptr()  # a bare capsule works too
    with pytest.raises(ValueError):
        _lib._ptr(_OnlyDLPack(a[:, ::2]))  # not C-contiguous
    with pytest.raises(TypeError):
        _lib._ptr(_OnlyDLPack(a.astype(np.float32)))  # wrong dtype
    cap = a.__dlpack__()
    _lib.from_dlpack_capsule(cap)
    _lib.from_dlpack_capsule(cap)  # reading does not consume the capsule


def test_tf_adapter_is_import_guarded():
    """TensorFlow is absent from this image: the adapter module imports, and says what it needs when used."""
    from multi_fidelity_gpflow_b200 import tf_adapter

    try:
        import tensorflow  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="TensorFlow"):
            tf_adapter.TFMultiFidelityGPR(__import__("numpy").zeros((4, 3)), __import__("numpy").zeros((4, 1)))

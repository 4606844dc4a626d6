"""The C-ABI library builds, loads, and exports every symbol include/ewvit.h declares.
No compute is launched here (no GPU in the CPU suite); argument validation that returns before
touching the device is exercised."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ewvit_build", os.path.join(REPO, "efficient-wavelet-vit_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def declared_symbols():
    text = open(os.path.join(REPO, "include", "ewvit.h")).read()
    return sorted(set(re.findall(r"EWVIT_API[^;(]*?\b(ewvit_\w+)\s*\(", text)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "ewvit_dwt3_haar_fwd" in syms and "ewvit_last_error" in syms and len(syms) >= 5


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(built_lib)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/ewvit.h but not exported"


def test_binding_covers_every_declared_symbol(built_lib):
    from ewvit import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.load().ewvit_abi_version() >= 1


def test_argument_validation_needs_no_device(built_lib):
    from ewvit import _lib
    L = _lib.load()
    # ragged size for the fused 3-level kernel -> EWVIT_ERR_UNSUPPORTED (-2) before any CUDA call
    assert L.ewvit_dwt3_haar_fwd(16, 1, 7, 8, 0, 0, 0, 0, 0, 0, None) == -2
    assert b"multiples of 8" in L.ewvit_last_error()
    # NULL input -> EWVIT_ERR_INVALID_ARG (-1)
    assert L.ewvit_dwt_haar_fwd(None, 1, 4, 4, None, None, None) == -1
    # empty input is a successful no-op
    assert L.ewvit_dwt_haar_fwd(None, 0, 4, 4, None, None, None) == 0
    assert L.ewvit_dwt3_haar_fwd(None, 0, 8, 8, 0, 0, 0, 0, 0, 0, None) == 0


def test_ops_refuse_cpu_tensors(built_lib):
    import torch
    from ewvit import EwvitError, ops
    with pytest.raises(EwvitError):
        ops.dwt_haar(torch.zeros(1, 3, 8, 8))
    with pytest.raises(EwvitError):
        ops.dwt3_haar(torch.zeros(1, 3, 8, 8))

"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol declared in
include/unigen_b200.h; argument validation works without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="unigen_b200.h", prefix="ug_"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built_lib):
    syms = _declared_symbols()
    assert "ug_decode_step_bf16" in syms and "ug_omni_attention_bf16" in syms
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in include/unigen_b200.h but not exported"
    from ml_unigen_b200 import _lib
    assert set(_lib.SYMBOLS) == set(syms)


def test_no_behaviour_switches_in_the_product_abi(built_lib):
    """Round-1 had thread-local ug_debug_set_* switches in the product header; test instrumentation is now explicit,
    per call, behind the ugt_ prefix of include/unigen_b200_test.h."""
    assert not [s for s in _declared_symbols() if "debug" in s]
    assert not hasattr(built_lib, "ug_debug_set_decode_mode")
    tsyms = _declared_symbols("unigen_b200_test.h", "ugt_")
    from ml_unigen_b200 import _lib
    assert set(tsyms) == set(_lib.TEST_SYMBOLS)
    for s in tsyms:
        assert hasattr(built_lib, s)


def test_abi_version_and_workspace(built_lib):
    assert built_lib.ug_abi_version() == 2
    assert built_lib.ug_decode_step_workspace_bytes(16, 1024, 8192) >= 16 * 1024 * 8
    assert built_lib.ug_decode_step_workspace_bytes(0, 1024, 8192) == 0


def test_argument_validation_without_gpu(built_lib):
    L = built_lib
    rc = L.ug_decode_step_bf16(None, None, 0.0, None, None, None, None, None, None, None, 0, 0, 0,
                               1, 16, 8192, 1, 1, 1.0, None, 0, None)
    assert rc == -1 and b"required" in L.ug_last_error()
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.cast(buf, ctypes.c_void_p).value
    p = (p + 255) & ~255
    rc = L.ug_decode_step_bf16(p, None, 0.0, p, p, p, None, None, None, None, 0, 0, 0,
                               1, 16, 1000, 1, 1, 1.0, p, 1 << 20, None)
    assert rc == -2 and b"V=1000" in L.ug_last_error()
    rc = L.ug_decode_step_bf16(p + 2, None, 0.0, p, p, p, None, None, None, None, 0, 0, 0,
                               1, 16, 8192, 1, 1, 1.0, p, 1 << 20, None)
    assert rc == -3
    rc = L.ug_decode_step_bf16(p, None, 0.0, p, p, p, None, None, None, None, 0, 0, 0,
                               1, 16, 8192, 1, 1, 1.0, p, 8, None)
    assert rc == -4
    rc = L.ug_omni_attention_bf16(p, p, p, p, p, p, 1, 128, 12, 2, 64, 1.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, None)
    assert rc == -2 and b"D=64" in L.ug_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from ml_unigen_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libunigen_b200.so")
    with pytest.raises(_lib.UnigenB200Error):
        _lib.lib()


def test_cpu_tensors_are_rejected(built_lib):
    import torch
    from ml_unigen_b200 import ops, sampling, _lib
    with pytest.raises(RuntimeError):
        sampling.mask_by_random_topk(torch.ones(1, 1), torch.rand(1, 8))
    with pytest.raises(_lib.UnigenB200Error):
        ops.mask_by_random_topk(torch.rand(1, 8), torch.ones(1, dtype=torch.int64), 1.0)

"""CPU: the C-ABI library loads and exports every symbol include/lf_engine.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from lumfuncmcmc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_if_needed():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def test_header_symbols_exported():
    _build_if_needed()
    header = open(os.path.join(ROOT, 'include', 'lf_engine.h')).read()
    declared = sorted(set(re.findall(r'\b(lf_[a-z0-9_]+)\s*\(', header)))
    assert sorted(_lib.EXPORTS) == declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_config_struct_layout_matches_header():
    # 8 int32 + 2 double + 5*2 double + 3 double
    assert ctypes.sizeof(_lib.LFConfig) == 8 * 4 + 8 * (2 + 10 + 3)


def test_version_and_error_strings():
    _build_if_needed()
    lib = _lib.load()
    assert b'sm_100a' in lib.lf_version()
    assert isinstance(lib.lf_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device context creation must fail loudly (never a silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _build_if_needed()
    lib = _lib.load()
    cfg = _lib.LFConfig()
    cfg.nfields, cfg.size_ln = 1, 11
    ctx = ctypes.c_void_p()
    rc = lib.lf_create(ctypes.byref(ctx), ctypes.byref(cfg))
    assert rc != 0
    assert b'no CPU fallback' in lib.lf_last_error()

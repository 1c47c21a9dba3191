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


def test_integration_md_binding_matches_the_header_struct():
    """The ctypes mirror of lf_config printed in INTEGRATION.md (the stub a maintainer would paste into the reference)
    has the same layout as the one the package itself uses."""
    import ctypes as C
    import os
    import re
    from lumfuncmcmc_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    m = re.search(r"class _LFConfig\(C\.Structure\):.*?\n\n", text, re.S)
    assert m, "INTEGRATION.md lost its _LFConfig stub"
    ns = {'C': C}
    exec(m.group(0), ns)
    doc = ns['_LFConfig']
    assert C.sizeof(doc) == C.sizeof(_lib.LFConfig)
    assert [(n, getattr(doc, n).offset, getattr(doc, n).size) for n, _ in doc._fields_] == \
           [(n, getattr(_lib.LFConfig, n).offset, getattr(_lib.LFConfig, n).size) for n, _ in _lib.LFConfig._fields_]
    # and every entry point the stub calls is exported
    for sym in set(re.findall(r"_lf\.(lf_\w+)", text)):
        assert sym in _lib.EXPORTS, sym

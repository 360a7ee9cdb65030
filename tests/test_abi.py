"""CPU: the C-ABI shared library loads and exports every symbol include/gnnrag_b200.h declares; argument
validation fails with status codes (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from gnn_rag_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gnnrag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(gr_[a-z0-9_]+)\s*\(", text))
    names.discard("gr_pad4")          # static inline helper
    return sorted(names)


def test_library_builds_and_loads():
    _build.build()
    lib = _lib.load()
    assert lib.gr_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_build.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for name in syms:
        assert hasattr(lib, name), "missing export: " + name
        assert name in _lib.SIGNATURES, "no ctypes signature for " + name
    assert sorted(_lib.SIGNATURES) == syms


def test_argument_validation_returns_status_codes():
    lib = _lib.load()
    rc = lib.gr_linear(None, 4, None, 4, None, None, 0, 0, None, 4, 4, 4, 4, 0, None)
    assert rc == -1 and b"null pointer" in lib.gr_last_error()
    rc = lib.gr_csr_build(None, None, None, 3, 0, 10, 5, None, None, None, None, None, None, None, None,
                          None, None, 0, None)
    assert rc == -1 and b"idx_bytes" in lib.gr_last_error()
    rc = lib.gr_set_option(b"no_such_option", 1)
    assert rc == -1
    assert lib.gr_set_option(b"agg_tma", 0) == 0
    assert lib.gr_csr_build_workspace_bytes(1000, 100) > 0
    assert lib.gr_rank_workspace_bytes(4, 100) == 4 * 100 * 8
    with pytest.raises(_lib.GrError):
        _lib.check(-1)


def test_models_refuse_cpu():
    import gnn_rag_b200 as G
    from gnn_rag_b200 import synthetic as S
    args = S.model_args("ReaRev", entity_dim=16, word_dim=8)
    m = G.ReaRev(args, 100, 10, 20)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(S.make_batch(0, 2, 10, 20, 100, 10, 20))
    with pytest.raises(NotImplementedError):
        m(S.make_batch(0, 2, 10, 20, 100, 10, 20), training=True)

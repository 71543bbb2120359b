"""The C-ABI library loads and exports every symbol include/mgb.h declares; config PODs agree.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

import oracle_lib as O
from metagraph_b200 import _lib
from metagraph_b200.config import mgb_config_t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mgb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    if not os.path.exists(_lib.DEFAULT_LIB):
        import __graft_entry__
        __graft_entry__.build()
    L = _lib.load_library()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(_lib.REQUIRED_SYMBOLS) == syms


def test_config_pod_layout():
    assert ctypes.sizeof(mgb_config_t) == O.lib().mgo_config_sizeof() == 16480
    assert mgb_config_t.score_matrix.offset == 96


def test_config_defaults_match_library():
    from metagraph_b200.config import cli_defaults, struct_defaults
    L = _lib.load_library()
    c = mgb_config_t()
    L.mgb_config_init(ctypes.byref(c))
    assert bytes(c) == bytes(struct_defaults().to_c())
    L.mgb_config_init_cli(ctypes.byref(c), 31, 0)
    assert bytes(c) == bytes(cli_defaults(31).to_c())
    L.mgb_config_init_cli(ctypes.byref(c), 11, 0)
    assert bytes(c) == bytes(cli_defaults(11).to_c())
    L.mgb_config_init_cli(ctypes.byref(c), 10, 1)                     # protein: BLOSUM62, forward strand only
    assert bytes(c) == bytes(cli_defaults(10, alphabet="protein").to_c())


def test_missing_library_fails_loudly():
    import pytest
    with pytest.raises(ImportError):
        _lib.load_library("/nonexistent/libmgb.so")

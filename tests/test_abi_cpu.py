"""The C-ABI library loads without a GPU and exports every symbol include/orl.h declares."""
import ctypes
import os
import re

import pytest

from tests import common
from openroborl_b200 import _abi


def _declared_symbols():
    text = open(os.path.join(common.ROOT, "include", "orl.h")).read()
    return sorted(set(re.findall(r"\b(orl_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _abi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _abi.SYMBOLS, "ctypes binding missing for " + name
    assert sorted(_abi.SYMBOLS) == declared
    assert lib.orl_abi_version() == _abi.ORL_ABI_VERSION


def test_struct_sizes_match_header():
    """Guard the ctypes mirrors against drift: sizes derived from the header's field lists."""
    assert ctypes.sizeof(_abi.OrlContactPoint) == 8 * 4
    assert ctypes.sizeof(_abi.OrlLegModel) == (27 + 9 + 3 + 9 + 18 + 9 + 15 + 3 + 1 + 2 + 6 + 1) * 4 + _abi.MAX_CP * 32
    assert ctypes.sizeof(_abi.OrlRobotModel) == (1 + 6 + 3 + 4 + 1) * 4 + 4 * ctypes.sizeof(_abi.OrlLegModel) + (1 + 32 + 32 + 96) * 4
    assert ctypes.sizeof(_abi.OrlStateLayout) == 30 * 4


def test_null_handle_calls_fail_cleanly():
    lib = _abi.load_library()
    assert lib.orl_destroy(None) == 0
    assert lib.orl_step(None, None, None, None, None, None) == _abi.ORL_ERR_INVALID
    assert lib.orl_create(None, None, None) == _abi.ORL_ERR_INVALID
    assert lib.orl_last_error(None) == b"null handle"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _abi.load_library()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(common.ROOT, "openroborl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "from tests" not in text, f

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


def test_struct_sizes_match_header(tmp_path):
    """Guard the ctypes mirrors against drift: sizeof() of every struct as the C compiler sees include/orl.h."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "orl.h"\nint main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(orl_config), '
                   'sizeof(orl_contact_point), sizeof(orl_leg_model), sizeof(orl_robot_model), sizeof(orl_state_layout)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(common.ROOT, "include"), "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirrors = [_abi.OrlConfig, _abi.OrlContactPoint, _abi.OrlLegModel, _abi.OrlRobotModel, _abi.OrlStateLayout]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


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

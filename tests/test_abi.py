"""The C-ABI library loads and exports exactly what include/deft_b200.h declares.  No compute, CPU only."""
import ctypes as C
import os
import re

import pytest

from deft_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "deft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(deft_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 13
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == names, "binding and header disagree"


def test_abi_version_and_struct_layout():
    assert _lib.lib.deft_b200_abi_version() == _lib.ABI_VERSION
    assert C.sizeof(_lib.Plan) == 144      # deft_plan_t: item/group layer (48 bytes) + unit layer (9 pointers + 6 ints)
    assert C.sizeof(_lib.Append) == 40


def test_argument_errors_are_reported_not_crashed():
    """Validation happens before any CUDA call, so this runs without a GPU."""
    rc = _lib.lib.deft_b200_flatten_fwd(None, 0, 0, None, None, 0, 0, 0, None, 0, 0, 1, 32, 8, 128, 128,
                                        None, 1, None, None, None, 1, None, None, None, None, 0, None)
    assert rc == -1 and "null" in _lib.last_error()
    with pytest.raises(_lib.DeftError):
        _lib.check(rc)
    rc = _lib.lib.deft_b200_kv_append(None, None, 0, 0, None, None, 0, 0, None, 1, 8, 128, None)
    assert rc == -1
    assert _lib.lib.deft_b200_build_tables(0, None, None, None, None, None, None, 0, 1, 128, 32, -1, 256, 8, 148, None, None) is None
    assert "tree" in _lib.last_error()


def test_workspace_size_is_monotone():
    f = _lib.lib.deft_b200_flatten_workspace_bytes
    a, b = f(64, 32, 8, 128, 2246, 81, None), f(64, 32, 8, 128, 4492, 162, None)
    assert b > a >= 81 * 8 * 128 * 128 * 2          # one fp16 partial tile per (block, kv-head) at least
    g = _lib.lib.deft_b200_node_workspace_bytes
    assert g(64, 32, 8, 128, 448, 128, 10208, None) > g(64, 32, 8, 128, 448, 128, 0, None)
    # geometries the tcgen05 kernel does not cover size the fp32 row layout of the warp-FMA path
    assert f(64, 32, 8, 32, 2246, 81, None) >= 81 * 32 * 32 * 32 * 4

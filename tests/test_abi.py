"""The C-ABI library loads and exports every symbol include/leibniz_b200.h declares (no GPU needed,
no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "leibniz_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    from leibnizgym_b200.build import build_library
    path = build_library()
    assert os.path.exists(path)
    from leibnizgym_b200 import _native
    lib = _native.load()
    assert lib.lg_version() == 111


def test_every_declared_symbol_is_exported_and_bound():
    from leibnizgym_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _native.SYMBOLS, f"{name} has no ctypes signature in _native.SYMBOLS"
    for name in _native.SYMBOLS:
        assert name in declared, f"{name} bound but not declared in the header"


def test_struct_layouts_agree():
    from leibnizgym_b200 import _native
    lib = _native.load()
    for which, struct in enumerate((_native.LgParams, _native.LgSimState, _native.LgBuffers, _native.LgControl,
                                    _native.LgRewardTerm, _native.LgHostStep)):
        assert lib.lg_struct_size(which) == ctypes.sizeof(struct), struct.__name__
    # the hot block of LgParams must stay inside the first two constant-cache lines
    assert _native.LgParams.stats_num_envs.offset + 8 <= 168


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only box."""
    from leibnizgym_b200 import _native
    lib = _native.load()
    assert lib.lg_scan_tiles(16384) == 128
    assert lib.lg_quat_mul(None, None, None, 4, None) == -1
    assert b"bad argument" in lib.lg_last_error()
    assert lib.lg_set_l2_fetch_granularity(48) == -1
    p, s, b = _native.LgParams(), _native.LgSimState(), _native.LgBuffers()
    assert lib.lg_post_physics(p, s, b, 0.0, None) == -1          # num_envs == 0
    p.num_envs, p.action_dim = 8, 7
    assert lib.lg_post_physics(p, s, b, 0.0, None) == -1          # bad action_dim
    assert b"action_dim" in lib.lg_last_error()
    with pytest.raises(ValueError):
        _native.check(-1, "x")

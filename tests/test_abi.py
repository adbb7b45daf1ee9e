"""The C-ABI library loads and exports exactly what include/pasnl.h declares (no compute calls: no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pasnl.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|const char \*)\s*\*?\s*(pasnl_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = 0 if args == ["void"] else len(args)
    return out


def test_header_declares_the_hot_path():
    d = declared()
    for name in ["pasnl_fps", "pasnl_gather_point", "pasnl_gather_point_grad", "pasnl_ball_query", "pasnl_group_point",
                 "pasnl_group_point_grad", "pasnl_knn", "pasnl_three_nn", "pasnl_three_interpolate",
                 "pasnl_three_interpolate_grad", "pasnl_as_core_fwd", "pasnl_as_core_bwd", "pasnl_as_resample_fwd",
                 "pasnl_as_resample_bwd", "pasnl_nl_core_fwd", "pasnl_nl_core_bwd", "pasnl_group_concat",
                 "pasnl_group_concat_grad", "pasnl_gather_rows", "pasnl_gather_rows_grad"]:
        assert name in d, name


def test_library_exports_every_declared_symbol_and_binding_table_matches():
    from pointasnl_b200 import _lib
    lib = _lib.load()
    d = declared()
    assert set(d) == set(_lib.SIGNATURES), set(d) ^ set(_lib.SIGNATURES)
    for name, nargs in d.items():
        assert hasattr(lib, name), f"{name} declared in pasnl.h but not exported"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args"
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    syms = {ln.split()[-1] for ln in exported.splitlines() if " T " in ln and "pasnl_" in ln}
    assert syms == set(d), syms ^ set(d)                         # nothing undeclared leaks out either


def test_version_and_error_strings_without_a_gpu():
    from pointasnl_b200 import _lib
    assert _lib.version().startswith("pasnl-b200") and "sm_100a" in _lib.version()
    lib = _lib.load()
    assert lib.pasnl_error_string(0) == b"ok"
    assert b"invalid argument" in lib.pasnl_error_string(-22)
    assert lib.pasnl_fps_workspace_bytes(2, 1024, 512) == 0
    assert lib.pasnl_fps_workspace_bytes(2, 70000, 64) == 2 * 70000 * 4
    assert lib.pasnl_nl_core_bwd_workspace_bytes(2, 100, 50, 32) == 2 * 100 * 4


def test_argument_validation_happens_before_any_launch():
    # bad shapes must come back as PASNL_EINVAL from the host-side checks (no device is touched)
    from pointasnl_b200 import _lib
    lib = _lib.load()
    assert lib.pasnl_fps(1, 0, 4, None, None, None, None) == -22
    assert lib.pasnl_knn(1, 8, 4, 9, None, None, None, None, None) == -22          # k > n
    assert lib.pasnl_knn(1, 800, 4, 129, None, None, None, None, None) == -22      # k > 128
    assert lib.pasnl_nl_core_fwd(1, 4, 4, 6, None, None, ctypes.c_float(1.0), None, None, None) == -22   # d % 4
    assert lib.pasnl_as_core_fwd(1, 33, 8, None, None, ctypes.c_float(1.0), None, None, None) == -22     # k > 32
    assert lib.pasnl_group_concat(1, 8, 0, 2, 4, None, None, None, 3, None, None, None) == -22           # stride < nsample
    assert lib.pasnl_fps(0, 8, 4, None, None, None, None) == 0                      # empty batch is a no-op


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from pointasnl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PasnlLibraryError, match="no fallback"):
        _lib.load()

"""The C-ABI library loads and exports every symbol include/rdfk.h declares
(no compute calls: there is no GPU on the CPU runner)."""
import ctypes
import os
import re

import pytest

from pmda_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "rdfk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rdfk_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = header_functions()
    assert "rdfk_rdf_hist" in names and "rdfk_rdf_sites" in names
    assert sorted(_cabi.EXPORTS) == names


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.rdfk_version() == 130


def test_argument_validation_needs_no_gpu():
    lib = _cabi.lib()
    # nbins < 1
    rc = lib.rdfk_rdf_hist(None, 1, 1, None, 1, None, 1, 0, 1, None, 0.0, 1.0,
                           0, None, 0, 0, None, None, 0, None)
    assert rc == -1
    assert b"nbins" in lib.rdfk_last_error()
    # r1 <= r0
    rc = lib.rdfk_rdf_hist(None, 1, 1, None, 1, None, 1, 0, 1, None, 2.0, 1.0,
                           10, None, 0, 0, None, None, 0, None)
    assert rc == -1
    # half an exclusion block
    rc = lib.rdfk_rdf_hist(None, 1, 1, None, 1, None, 1, 0, 1, None, 0.0, 1.0,
                           10, None, 3, 0, None, None, 0, None)
    assert rc == -1
    # nothing to do is not an error
    rc = lib.rdfk_rdf_hist(None, 0, 1, None, 1, None, 1, 0, 1, None, 0.0, 1.0,
                           10, None, 0, 0, None, None, 0, None)
    assert rc == 0
    # null pointers with work to do
    rc = lib.rdfk_rdf_hist(None, 1, 1, None, 1, None, 1, 0, 1, None, 0.0, 1.0,
                           10, None, 0, 0, None, None, 0, None)
    assert rc == -1
    rc = lib.rdfk_rdf_sites(None, 1, 1, None, None, None, None, None, 1, 4, 7,
                            None, 0.0, 1.0, 10, None, None, None)
    assert rc == -1
    with pytest.raises(_cabi.RdfkError):
        _cabi.check(rc)
    # rdfk_hbonds: cut-off, missing total, missing buffers
    hb = lib.rdfk_hbonds
    assert hb(None, 1, 3, None, None, 1, None, 1, 1, None, 0.0, 1.0, 150.0, 0,
              None, None, None, None, None, None, 0, None) == -1
    assert b"max_cutoff" in lib.rdfk_last_error()
    assert hb(None, 1, 3, None, None, 1, None, 1, 1, None, 3.0, 1.0, 150.0, 0,
              None, None, None, None, None, None, 0, None) == -1
    assert hb(None, 1, 3, None, None, 1, None, 1, 7, None, 3.0, 1.0, 150.0, 0,
              None, None, None, None, None, None, 0, None) == -1
    assert lib.rdfk_hbonds_workspace_bytes(4, 600, 300) > 4 * 900 * 12
    assert lib.rdfk_hbonds_workspace_bytes(-1, 1, 1) == 0


def test_workspace_query():
    a = _cabi.rdf_workspace_bytes(4, 3000, 3000, 75, True)
    b = _cabi.rdf_workspace_bytes(4, 3000, 3000, 75, False)
    assert 0 < a < b
    assert a % 256 == 0 and b % 256 == 0
    # 3 float planes per padded atom per frame dominate
    assert a >= 4 * 3 * 3072 * 4


def test_missing_library_is_loud(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/librdfk.so")
    with pytest.raises(_cabi.RdfkError, match="no CPU fallback"):
        _cabi.lib()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pmda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src, f


def test_debug_switches_are_an_explicit_call_not_an_environment_variable():
    """the release library read RDFK_DEBUG with getenv on every
    call, so a stray export could switch off pruning in production."""
    csrc = os.path.join(ROOT, "pmda_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh", ".inl")):
            src = open(os.path.join(csrc, f)).read()
            assert "getenv" not in src, f
    blob = open(_cabi.LIB_PATH, "rb").read()
    assert b"RDFK_DEBUG" not in blob
    assert _cabi.set_debug_flags(5) == 0
    assert _cabi.set_debug_flags(0) == 5
    assert not _cabi.build_has_checks()


def test_check_build_exports_the_same_symbols():
    from pmda_b200.csrc import build as b
    path = b.build_check()
    lib = ctypes.CDLL(path)
    for name in header_functions():
        assert hasattr(lib, name), name
    lib.rdfk_build_has_checks.restype = ctypes.c_int
    assert lib.rdfk_build_has_checks() == 1

"""CPU-side checks of the drop-in boundary: the shared library loads and exports every
symbol that include/lineax_b200.h declares (no compute calls -- those are `-m gpu`)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lineax_b200.h")
LIB = os.path.join(ROOT, "lineax_b200", "liblineax_b200.so")


def _declared_symbols():
    pre = subprocess.run(["gcc", "-E", "-P", HEADER], check=True, capture_output=True, text=True).stdout
    return sorted(set(re.findall(r"\b(lxb_[a-z0-9_]+)\s*\(", pre)))


def test_header_is_plain_c():
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(LIB)
    syms = _declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_covers_header():
    import lineax_b200._native as nat

    unbound = [s for s in _declared_symbols()
               if s not in nat._SIGS and s not in ("lxb_version", "lxb_error_string", "lxb_launch_count")]
    assert not unbound, f"no ctypes signature for: {unbound}"
    assert nat.lib.lxb_version() >= 100
    assert b"bad argument" in nat.lib.lxb_error_string(-1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lineax_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lineax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle." not in src.replace("oracle/getf2.c", ""), f

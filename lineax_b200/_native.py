"""ctypes binding of the C ABI declared in `include/lineax_b200.h`.

There is NO fallback: if `liblineax_b200.so` is missing, importing this module
raises, and every wrapper raises on a non-zero return code.  Build the library
with `python -c "import __graft_entry__ as g; g.build()"` (or `make -C
lineax_b200/csrc`).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblineax_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"lineax_b200: native library {LIB_PATH} not found. It is required (there is no "
        "CPU/PyTorch fallback). Build it with `make -C lineax_b200/csrc` or "
        "`python -c 'import __graft_entry__ as g; g.build()'`."
    )

lib = ctypes.CDLL(LIB_PATH)

# flags (include/lineax_b200.h)
TRANS = 1 << 0
NSD = 1 << 1
MAXSTEPS_GIVEN = 1 << 2
X64_BREAKDOWN = 1 << 3
HAS_Y0 = 1 << 4
UNIT_DIAG = 1 << 5
LOWER = 1 << 6
QT_ONLY = 1 << 7

E_BADARG, E_UNSUPPORTED, E_WORKSPACE, E_ALIGN = -1, -2, -3, -4

lib.lxb_version.restype = ctypes.c_int
lib.lxb_error_string.restype = ctypes.c_char_p
lib.lxb_error_string.argtypes = [ctypes.c_int]
lib.lxb_launch_count.restype = c_int64

P = c_void_p
_SIGS = {}


def _decl(name, argtypes, restype=ctypes.c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    _SIGS[name] = fn
    return fn


for _sfx, _T in (("f32", c_float), ("f64", c_double)):
    _decl(f"lxb_lu_factor_{_sfx}", [P, c_int64, P, P, c_int64, c_int32, P])
    _decl(f"lxb_lu_solve_{_sfx}", [P, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int32, c_int32, P])
    _decl(f"lxb_lu_factor_solve_{_sfx}", [P, c_int64, P, c_int64, P, P, P, c_int64, c_int32, P])
    _decl(
        f"lxb_cg_{_sfx}",
        [P, c_int64, P, c_int64, P, c_int64, P, P, P, c_int64, c_int32, _T, _T, c_int32, c_int32,
         c_int32, P, c_size_t, P],
    )
    _decl(f"lxb_cg_workspace_{_sfx}", [c_int64, c_int32], c_size_t)
    _decl(f"lxb_postprocess_{_sfx}", [P, c_int64, c_int32, P, c_int64, c_int32, P, c_int64, P])
    _decl(f"lxb_matvec_{_sfx}", [P, c_int64, P, c_int64, P, c_int64, c_int32, c_int32, c_int32, P])
    _decl(f"lxb_diag_mv_{_sfx}", [P, c_int64, P, c_int64, P, c_int64, c_int32, P])
    _decl(f"lxb_tridiag_mv_{_sfx}", [P, P, P, c_int64, P, c_int64, P, c_int64, c_int32, P])
    _decl(f"lxb_norms_{_sfx}", [P, c_int64, P, c_int64, P, c_int64, c_int64, P])
    _decl(f"lxb_lu_solve_multi_{_sfx}", [P, c_int64, P, c_int64, P, P, c_int64, c_int32, c_int32, c_int32, P])
    _decl(f"lxb_cholesky_solve_multi_{_sfx}", [P, c_int64, P, P, c_int64, c_int32, c_int32, c_int32, P])
    _decl(f"lxb_triangular_solve_multi_{_sfx}", [P, c_int64, P, P, c_int64, c_int32, c_int32, c_int32, P])
    _decl(f"lxb_gram_{_sfx}", [P, c_int64, P, c_int64, c_int32, c_int32, c_int32, P])
    _decl(f"lxb_cholesky_factor_{_sfx}", [P, c_int64, P, c_int64, c_int32, c_int32, P])
    _decl(f"lxb_cholesky_solve_{_sfx}", [P, c_int64, P, c_int64, P, c_int64, c_int32, c_int32, P])
    _decl(f"lxb_qr_factor_{_sfx}", [P, c_int64, P, P, c_int64, c_int32, c_int32, P, c_size_t, P])
    _decl(f"lxb_qr_factor_workspace_{_sfx}", [c_int64, c_int32, c_int32], c_size_t)
    _decl(f"lxb_qr_solve_{_sfx}",
          [P, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int32, c_int32, c_int32, P, c_size_t, P])
    _decl(f"lxb_qr_solve_workspace_{_sfx}", [c_int64, c_int32, c_int32], c_size_t)
    _decl(f"lxb_tridiagonal_solve_{_sfx}",
          [P, P, P, c_int64, P, c_int64, P, c_int64, c_int32, P, c_size_t, P])
    _decl(f"lxb_tridiagonal_workspace_{_sfx}", [c_int64, c_int32], c_size_t)
    _decl(f"lxb_diagonal_solve_{_sfx}", [P, c_int64, P, c_int64, P, c_int64, c_int32, _T, P])
    _decl(f"lxb_triangular_solve_{_sfx}", [P, c_int64, P, c_int64, P, c_int64, c_int32, c_int32, P])
    _decl(
        f"lxb_bicgstab_{_sfx}",
        [P, c_int64, P, c_int64, P, c_int64, P, P, P, c_int64, c_int32, _T, _T, c_int32, c_int32, P,
         c_size_t, P],
    )
    _decl(f"lxb_bicgstab_workspace_{_sfx}", [c_int64, c_int32], c_size_t)
    _decl(
        f"lxb_gmres_{_sfx}",
        [P, c_int64, P, c_int64, P, c_int64, P, P, P, c_int64, c_int32, _T, _T, c_int32, c_int32,
         c_int32, c_int32, P, c_size_t, P],
    )
    _decl(f"lxb_gmres_workspace_{_sfx}", [c_int64, c_int32, c_int32], c_size_t)
    _decl(
        f"lxb_lsmr_{_sfx}",
        [P, c_int64, P, c_int64, P, P, P, P, c_int64, c_int32, c_int32, _T, _T, _T, c_int64, c_int32,
         P, c_size_t, P],
    )
    _decl(f"lxb_lsmr_workspace_{_sfx}", [c_int64, c_int32, c_int32], c_size_t)
    _decl(
        f"lxb_gmres_rowsharded_{_sfx}",
        [P, P, P, P, P, c_int32, c_int32, c_int32, _T, _T, c_int32, c_int32, c_int32, c_int32, P, c_size_t,
         P, c_int32, c_int32, P],
    )
    _decl(f"lxb_gmres_rowsharded_workspace_{_sfx}", [c_int32, c_int32], c_size_t)
    _decl(
        f"lxb_cg_rowsharded_{_sfx}",
        [P, P, P, P, P, c_int32, c_int32, c_int32, _T, _T, c_int32, c_int32, c_int32, P, c_size_t, P, c_int32,
         c_int32, P],
    )
    _decl(f"lxb_cg_rowsharded_workspace_{_sfx}", [c_int32], c_size_t)
    _decl(
        f"lxb_bicgstab_rowsharded_{_sfx}",
        [P, P, P, P, P, c_int32, c_int32, c_int32, _T, _T, c_int32, c_int32, P, c_size_t, P, c_int32, c_int32, P],
    )
    _decl(f"lxb_gmres_rowsharded_symm_bytes_{_sfx}", [c_int32], c_size_t)
    _decl(
        f"lxb_lsmr_rowsharded_{_sfx}",
        [P, P, P, P, P, P, c_int32, c_int32, c_int32, _T, _T, _T, c_int64, c_int32, P, c_size_t, P, c_int32,
         c_int32, P],
    )
    _decl(f"lxb_lsmr_rowsharded_workspace_{_sfx}", [c_int32, c_int32], c_size_t)
    _decl(f"lxb_lsmr_rowsharded_symm_bytes_{_sfx}", [c_int32, c_int32], c_size_t)
_decl("lxb_lu_factor_solve_f32_host", [P, P, P, c_int64, c_int32, P, c_size_t, P])
_decl("lxb_host_scratch_bytes", [c_int64, c_int32, c_int32], c_size_t)
_decl("lxb_fp32_fma_probe", [P, c_int32, c_int32, P, P])


class NativeError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = lib.lxb_error_string(rc).decode()
        raise NativeError(f"{what} failed with code {rc}: {msg}")


def call(name: str, *args):
    """Call an int-returning entry point and raise on failure."""
    check(_SIGS[name](*args), name)


def fn(name: str):
    return _SIGS[name]


def launch_count() -> int:
    return int(lib.lxb_launch_count())


def suffix(dtype) -> str:
    import torch

    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(
        f"lineax_b200 native kernels support float32 and float64 operands, got {dtype}. "
        "(Complex dtypes are outside the accelerated hot path.)"
    )

"""CPU oracle for the lineax solve hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the algorithms in the reference
(`/root/reference/lineax/_solver/*.py`, `_solve.py:97-129`).  It exists so the
CUDA path can be checked for parity; it is NOT part of the product.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it.  `lineax_b200/` never imports `oracle`.

Pinning status
--------------
The reference is pure Python on top of JAX/jaxlib/Equinox, none of which are
installable in this image (no network), so the reference itself cannot be run
here.  The oracle is therefore pinned against
  * the closed-form / hard-coded vectors held by the reference's own tests
    (tests/test_solve.py:41-61, tests/test_singular.py:75-99,
    tests/test_lsmr.py:7-30, tests/test_solve.py:174-194,
    tests/test_adjoint.py:84-130) -- see tests/test_oracle_golden.py, and
  * the same property the reference test-suite uses everywhere else
    (agreement with numpy.linalg.solve / lstsq at 1e-10, tests/test_well_posed.py:31-52).
Quantities the reference's tests never assert (LU pivots, iteration counts,
fp32 behaviour) are "parity unpinned": for those this restatement (and, for
pivots, LAPACK getrf via SciPy) is the only arbiter.

Third-party arithmetic restated here: jax/jaxlib (`jax>=0.10.0`,
pyproject.toml:50, no lock file) -- lu_factor/lu_solve, cho_factor/cho_solve,
qr(mode="raw")/ormqr/solve_triangular, tridiagonal_solve lower to LAPACK
getrf/getrs, potrf/potrs, geqrf/ormqr/trtrs and gtsv on the CPU backend; the
oracle calls the same LAPACK routines through scipy.linalg.lapack (OpenBLAS).
"""
from .results import RESULTS  # noqa: F401
from .krylov import cg, bicgstab, gmres, lsmr  # noqa: F401
from .direct import (  # noqa: F401
    lu_init, lu_compute, cholesky_init, cholesky_compute, qr_init, qr_compute,
    tridiagonal_compute, diagonal_compute, triangular_compute,
)
from .solve import postprocess, max_norm, two_norm, tree_dot, resolve_rcond  # noqa: F401
from . import gen  # noqa: F401

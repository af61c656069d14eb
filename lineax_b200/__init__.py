"""lineax_b200: B200-native drop-in for the solve hot path of patrick-kidger/lineax.

Public names mirror `lineax/__init__.py:17-84` for the part of the library on the hot
path: `linear_solve`, the solvers, the materialised operators, tags,
`RESULTS`, `Solution`.  All numerical work runs in hand-written sm_100a CUDA kernels
behind the C ABI in `include/lineax_b200.h`; there is no CPU fallback.
"""
from . import _native as _native  # fails loudly when the CUDA library is missing
from ._norm import max_norm as max_norm, rms_norm as rms_norm, tree_dot as tree_dot, two_norm as two_norm
from ._operator import (
    AbstractLinearOperator as AbstractLinearOperator,
    DiagonalLinearOperator as DiagonalLinearOperator,
    IdentityLinearOperator as IdentityLinearOperator,
    MatrixLinearOperator as MatrixLinearOperator,
    PyTreeLinearOperator as PyTreeLinearOperator,
    TaggedLinearOperator as TaggedLinearOperator,
    TridiagonalLinearOperator as TridiagonalLinearOperator,
    conj as conj,
    diagonal as diagonal,
    has_unit_diagonal as has_unit_diagonal,
    is_diagonal as is_diagonal,
    is_lower_triangular as is_lower_triangular,
    is_negative_semidefinite as is_negative_semidefinite,
    is_positive_semidefinite as is_positive_semidefinite,
    is_symmetric as is_symmetric,
    is_tridiagonal as is_tridiagonal,
    is_upper_triangular as is_upper_triangular,
    linearise as linearise,
    materialise as materialise,
    tridiagonal as tridiagonal,
)
from ._solution import RESULTS as RESULTS, LinearSolveError as LinearSolveError, Solution as Solution
from ._solve import (
    AbstractLinearSolver as AbstractLinearSolver,
    AutoLinearSolver as AutoLinearSolver,
    config as config,
    invert as invert,
    linear_solve as linear_solve,
)
from ._solver import *  # noqa: F401,F403
from ._tags import (
    diagonal_tag as diagonal_tag,
    lower_triangular_tag as lower_triangular_tag,
    negative_semidefinite_tag as negative_semidefinite_tag,
    positive_semidefinite_tag as positive_semidefinite_tag,
    symmetric_tag as symmetric_tag,
    transpose_tags as transpose_tags,
    tridiagonal_tag as tridiagonal_tag,
    unit_diagonal_tag as unit_diagonal_tag,
    upper_triangular_tag as upper_triangular_tag,
)
from ._tree import ShapeDtypeStruct as ShapeDtypeStruct, set_default_device as set_default_device

__version__ = "0.1.0"

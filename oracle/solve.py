"""Vector kernels and result post-processing (test infrastructure, see oracle/__init__.py).

Follows lineax/_norm.py:27-139, lineax/_misc.py:30-38 and lineax/_solve.py:104-123.
"""
import numpy as np

from .results import RESULTS


def tree_dot(x, y):
    # lineax/_norm.py:27-47 (real dtypes: conj is the identity)
    return np.dot(x.reshape(-1), y.reshape(-1))


def two_norm(x):
    # lineax/_norm.py:59-82: size-1 shortcut |x|, otherwise sqrt(sum of squares)
    x = np.asarray(x)
    if x.size == 1:
        return np.abs(x.reshape(()))
    return np.sqrt(tree_dot(x, x))


def max_norm(x):
    # lineax/_norm.py:123-139; NaN-propagating like jnp.max; size 0 -> 0
    x = np.asarray(x)
    if x.size == 0:
        return x.dtype.type(0.0)
    return np.max(np.abs(x))


def resolve_rcond(rcond, n, m, dtype):
    # lineax/_misc.py:30-38
    eps = np.finfo(dtype).eps
    if rcond is None:
        return dtype.type(2 * eps * max(n, m))
    return dtype.type(eps) if rcond < 0 else dtype.type(rcond)


def postprocess(solution, result, vector):
    """lineax/_solve.py:104-123: successful+nonfinite x -> singular; singular+nonfinite b -> nonfinite_input."""
    if result == RESULTS.successful and not np.all(np.isfinite(solution)):
        result = RESULTS.singular
    if result == RESULTS.singular and not np.all(np.isfinite(vector)):
        result = RESULTS.nonfinite_input
    return result

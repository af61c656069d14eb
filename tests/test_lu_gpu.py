"""LU parity (GPU): CUDA kernels vs the C oracle (bit-exact) and vs LAPACK getrf (pivots).
Reference: lineax/_solver/lu.py:43-66."""
import numpy as np
import pytest
import torch

from oracle import clib, gen, lu_init, lu_compute
from tests.helpers import assert_close, dev, host

pytestmark = pytest.mark.gpu


def _ops():
    from lineax_b200 import _ops

    return _ops


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 13, 16, 20, 31, 32])
def test_factor_solve_bit_exact_vs_c_oracle(n, dtype):
    a, b, _ = gen.gaussian_systems(100 + n, 257, n, dtype)
    x_ref, lu_ref, piv_ref = clib.lu_factor_solve(a, b)
    x, lu, piv = _ops().lu_factor_solve(dev(a), dev(b), True)
    assert np.array_equal(host(piv), piv_ref), "pivot indices must be bit-exact"
    assert np.array_equal(host(lu), lu_ref), "LU factors must be bit-exact"
    assert np.array_equal(host(x), x_ref), "fused solve must be bit-exact"
    # state-less fused path gives the same x
    x2, _, _ = _ops().lu_factor_solve(dev(a), dev(b), False)
    assert np.array_equal(host(x2), x_ref)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [33, 40, 64, 100, 200, 260])
def test_block_tier_bit_exact_vs_c_oracle(n, dtype):
    a, b, _ = gen.gaussian_systems(200 + n, 5, n, dtype)
    x_ref, lu_ref, piv_ref = clib.lu_factor_solve(a, b)
    x, lu, piv = _ops().lu_factor_solve(dev(a), dev(b), True)
    assert np.array_equal(host(piv), piv_ref)
    assert np.array_equal(host(lu), lu_ref)
    assert np.array_equal(host(x), x_ref)
    lu2, piv2 = _ops().lu_factor(dev(a))
    assert np.array_equal(host(piv2), piv_ref) and np.array_equal(host(lu2), lu_ref)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [4, 16, 32, 48, 130])
@pytest.mark.parametrize("trans", [False, True])
def test_factor_then_solve(n, dtype, trans):
    a, b, _ = gen.gaussian_systems(300 + n, 65, n, dtype)
    lu_ref, piv_ref = clib.lu_factor(a)
    x_ref = clib.lu_solve(lu_ref, piv_ref, b, trans=int(trans))
    lu, piv = _ops().lu_factor(dev(a))
    assert np.array_equal(host(piv), piv_ref)
    assert np.array_equal(host(lu), lu_ref)
    x = _ops().lu_solve(lu, piv, dev(b), trans)
    assert np.array_equal(host(x), x_ref)
    # and against LAPACK getrs on the same factors (different rounding order): tolerance
    for i in range(0, 65, 16):
        x_lapack = lu_compute(lu_init(a[i]), b[i], trans=int(trans))
        assert_close(host(x)[i], x_lapack, dtype, factor=1e3 * n)


def test_pivots_match_lapack_getrf_c2_sample():
    """C2 inputs (65536 x 32^2 f32): pivots equal LAPACK's on a 4096-system sample."""
    a, b, _ = gen.gaussian_systems(1, 4096, 32, np.float32)
    _, _, piv = _ops().lu_factor_solve(dev(a), dev(b), True)
    piv = host(piv)
    mism = sum(not np.array_equal(lu_init(a[i])[1], piv[i]) for i in range(4096))
    assert mism == 0, f"{mism} of 4096 systems differ from LAPACK getrf pivots"


def test_c2_full_size_properties():
    """Full C2 size through size-independent properties: residual + permutation validity."""
    a, b, _ = gen.gaussian_systems(1, 65536, 32, np.float32)
    A, B = dev(a), dev(b)
    x, lu, piv = _ops().lu_factor_solve(A, B, True)
    piv_h = host(piv)
    assert piv_h.min() >= 0 and piv_h.max() < 32
    assert np.all(piv_h >= np.arange(32)[None, :]), "getrf pivots satisfy piv[k] >= k"
    x_ref, _, piv_ref = clib.lu_factor_solve(a, b)
    assert np.array_equal(piv_h, piv_ref)
    assert np.array_equal(host(x), x_ref)
    r = np.einsum("bij,bj->bi", a.astype(np.float64), host(x).astype(np.float64)) - b
    scale = np.abs(a).sum(-1).max(-1) * np.abs(host(x)).max(-1) + np.abs(b).max(-1)
    assert np.max(np.abs(r).max(-1) / scale) < 32 * 1.2e-7 * 50


def test_ties_follow_isamax_first_index():
    """Exact ties in |a_ik| (integer / structured matrices) pick LAPACK's first index."""
    rng = np.random.default_rng(7)
    a = rng.integers(-2, 3, size=(512, 32, 32)).astype(np.float32)
    a += 0  # many ties and zero pivots
    lu_ref, piv_ref = clib.lu_factor(a)
    _, piv = _ops().lu_factor(dev(a))
    assert np.array_equal(host(piv), piv_ref)
    pois = np.stack([gen.poisson_matrix(32, np.float32)] * 3)
    _, piv = _ops().lu_factor(dev(pois))
    assert np.array_equal(host(piv)[0], lu_init(pois[0])[1])


def test_broadcast_operands():
    """vmap(in_axes=(None, 0)): one matrix, many right-hand sides (stride-0 operand)."""
    a, b, _ = gen.gaussian_systems(5, 1, 32, np.float32)
    bs = np.random.default_rng(0).standard_normal((100, 32)).astype(np.float32)
    x, _, _ = _ops().lu_factor_solve(dev(a[0]), dev(bs), False)
    x_ref, _, _ = clib.lu_factor_solve(np.repeat(a, 100, 0), bs)
    assert np.array_equal(host(x), x_ref)


def test_singular_gives_nonfinite():
    a = np.zeros((2, 4, 4), np.float32)
    a[1] = np.eye(4)
    x, _, _ = _ops().lu_factor_solve(dev(a), dev(np.ones((2, 4), np.float32)), False)
    x = host(x)
    assert not np.all(np.isfinite(x[0])) and np.allclose(x[1], 1.0)


@pytest.mark.parametrize("scale", [1e-42, 1e-38, 1e-30, 1e30, 3e37])
def test_extreme_pivots_take_the_ieee_division_path(scale):
    """Denormal / huge pivots: 1/pivot leaves the fast MUFU.RCP+Newton range and must still be the
    correctly rounded IEEE quotient the C oracle computes (rcp_slow in csrc/lu_tma.cu)."""
    a, b, _ = gen.gaussian_systems(77, 130, 32, np.float32)
    a = (a.astype(np.float64) * scale).astype(np.float32)
    a[3, 5, :] = 0.0  # one singular system: inf / NaN must propagate identically
    x_ref, lu_ref, piv_ref = clib.lu_factor_solve(a, b)
    x, lu, piv = _ops().lu_factor_solve(dev(a), dev(b), True)
    assert np.array_equal(host(piv), piv_ref)
    assert np.array_equal(host(lu), lu_ref, equal_nan=True)
    assert np.array_equal(host(x), x_ref, equal_nan=True)
    x2, _, _ = _ops().lu_factor_solve(dev(a), dev(b), False)
    assert np.array_equal(host(x2), x_ref, equal_nan=True)


def test_strided_and_odd_batches_tma_path():
    """Batch sizes around the 8-warp CTA granularity and a strided operand (systems 2 apart)."""
    for batch in (1, 7, 8, 9, 1185, 1191):
        a, b, _ = gen.gaussian_systems(batch, batch, 32, np.float32)
        x_ref, lu_ref, piv_ref = clib.lu_factor_solve(a, b)
        x, lu, piv = _ops().lu_factor_solve(dev(a), dev(b), True)
        assert np.array_equal(host(x), x_ref) and np.array_equal(host(lu), lu_ref)
        assert np.array_equal(host(piv), piv_ref)
        x2, _, _ = _ops().lu_factor_solve(dev(a), dev(b), False)
        assert np.array_equal(host(x2), x_ref)
    a, b, _ = gen.gaussian_systems(3, 64, 32, np.float32)
    x_ref, _, _ = clib.lu_factor_solve(a[::2], b[::2])
    x, _, _ = _ops().lu_factor_solve(dev(a)[::2], dev(b)[::2], False)
    assert np.array_equal(host(x), x_ref)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [5, 32, 100, 300, 700])
@pytest.mark.parametrize("trans", [False, True])
def test_multi_rhs_bit_identical_to_single_vector_solves(n, dtype, trans):
    """`state=` reuse / `lx.invert` / vmap(in_axes=(None, 0)) (lineax/_solve.py:732-740, 809-871): many
    vectors against ONE factorisation run the multi-RHS kernels (csrc/multi_rhs.cu) and must equal the
    C oracle's getrs on every vector bit for bit."""
    a, _, _ = gen.gaussian_systems(400 + n, 1, n, dtype)
    rng = np.random.default_rng(n)
    nrhs = 77
    bs = rng.standard_normal((nrhs, n)).astype(dtype)
    lu_ref, piv_ref = clib.lu_factor(a)
    x_ref = clib.lu_solve(np.repeat(lu_ref, nrhs, 0), np.repeat(piv_ref, nrhs, 0), bs, trans=int(trans))
    lu, piv = _ops().lu_factor(dev(a))
    x = _ops().lu_solve(lu[0], piv[0], dev(bs), trans)  # unbatched factors, batched vectors
    assert np.array_equal(host(x), x_ref)


def test_multi_rhs_through_linear_solve_state_and_vmap():
    """The public path: one `solver.init`, vmapped `linear_solve(..., state=state)` over 200 vectors."""
    import lineax_b200 as lx

    rng = np.random.default_rng(3)
    n, nrhs = 64, 200
    a = rng.standard_normal((n, n)) + 8 * np.eye(n)
    bs = rng.standard_normal((nrhs, n))
    A, Bs = dev(a), dev(bs)
    for solver, op in ((lx.LU(), lx.MatrixLinearOperator(A)),
                       (lx.Cholesky(), lx.MatrixLinearOperator(dev(a @ a.T), lx.positive_semidefinite_tag)),
                       (lx.Triangular(), lx.MatrixLinearOperator(dev(np.triu(a)), lx.upper_triangular_tag))):
        state = solver.init(op, {})
        before = lx._native.launch_count()
        xs = torch.func.vmap(lambda v: lx.linear_solve(op, v, solver, state=state, throw=False).value)(Bs)
        torch.cuda.synchronize()
        mat = op.as_matrix().cpu().numpy()
        ref = np.linalg.solve(mat, bs.T).T
        assert np.max(np.abs(host(xs) - ref)) / np.max(np.abs(ref)) < 1e-9, type(solver).__name__
        assert lx._native.launch_count() - before <= 4, "one launch for all vectors, not one per vector"

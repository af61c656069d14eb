/*
 * lineax_b200 -- C ABI of the B200-native solve hot path.
 *
 * The reference (patrick-kidger/lineax) has no FFI today: its seam is the Python
 * ABC `AbstractLinearSolver` (lineax/_solve.py:343-480) whose `init` / `compute`
 * are invoked at lineax/_solve.py:790 and lineax/_solve.py:98.  Every entry point
 * below replaces the body of one of those methods for materialised operators and
 * is what an XLA-FFI (`jax.ffi`) handler or the ctypes binding in
 * lineax_b200/_native.py forwards to 1:1.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in `_host`.
 *  - Matrices are row-major; `batch` independent systems are laid out with an
 *    element stride between consecutive systems (`stride_*`, 0 = broadcast the
 *    same operand to every system -- the vmap(in_axes=None) case).
 *  - The caller owns every buffer; the library never allocates or frees device
 *    memory and keeps no global state.  Calls are stream-ordered and never
 *    synchronise the device.
 *  - Return value: 0 ok; <0 bad argument (LXB_E_*); >0 a cudaError_t.
 *    Numerical failure is DATA (`result[]` holds lineax RESULTS codes,
 *    lineax/_solution.py:52-68), never a return code.
 *  - `result[]` codes are the solver's own verdict (what `compute` returns);
 *    the non-finite rewriting of lineax/_solve.py:104-123 is applied by
 *    lxb_postprocess_* (or fused, where stated).
 */
#ifndef LINEAX_B200_H_
#define LINEAX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* lxb_stream_t; /* == cudaStream_t */

/* RESULTS codes, lineax/_solution.py:52-68 (definition order). */
enum {
  LXB_SUCCESSFUL = 0,
  LXB_MAX_STEPS_REACHED = 1,
  LXB_SINGULAR = 2,
  LXB_BREAKDOWN = 3,
  LXB_STAGNATION = 4,
  LXB_CONLIM = 5,
  LXB_NONFINITE_INPUT = 6
};

/* Argument errors. */
enum {
  LXB_E_BADARG = -1,      /* null pointer / negative size */
  LXB_E_UNSUPPORTED = -2, /* shape outside what the kernels cover */
  LXB_E_WORKSPACE = -3,   /* workspace too small */
  LXB_E_ALIGN = -4        /* pointer/stride alignment requirement violated */
};

/* flags */
enum {
  LXB_TRANS = 1 << 0,          /* solve with the transposed operator (lu.py:62, qr.py:74) */
  LXB_NSD = 1 << 1,            /* operator is negative definite: solve with -A, negate x (cg.py:100,224) */
  LXB_MAXSTEPS_GIVEN = 1 << 2, /* max_steps was not None (selects max_steps_reached vs singular, cg.py:213-222) */
  LXB_X64_BREAKDOWN = 1 << 3,  /* BiCGStab: `== 0` breakdown test (jax_enable_x64), bicgstab.py:110-113 */
  LXB_HAS_Y0 = 1 << 4,         /* x holds the initial guess on entry (options["y0"]) */
  LXB_UNIT_DIAG = 1 << 5,      /* triangular solve: unit diagonal */
  LXB_LOWER = 1 << 6,          /* triangular solve: lower triangular */
  LXB_QT_ONLY = 1 << 7         /* qr_solve: return (Q^T b)[:cols] without the R solve (row-sharded TSQR) */
};

int lxb_version(void);
const char* lxb_error_string(int code);
/* Number of kernels this library has launched in this process (bench `gpu_launches`). */
int64_t lxb_launch_count(void);

/* ------------------------------------------------------------------ LU --
 * lineax/_solver/lu.py:43-66.  `piv` is LAPACK getrf's row-swap sequence,
 * 0-based int32 (jax.scipy.linalg.lu_factor convention).
 * factor      : init()      A[batch,n,n] -> lu[batch,n,n], piv[batch,n]
 * solve       : compute()   x = lu_solve((lu,piv), b, trans)
 * factor_solve: init+compute fused, A is read once; lu/piv may be NULL (state not kept).
 */
#define LXB_DECL_LU(sfx, T)                                                                      \
  int lxb_lu_factor_##sfx(const T* A, int64_t stride_A, T* lu, int32_t* piv, int64_t batch,      \
                          int32_t n, lxb_stream_t stream);                                       \
  int lxb_lu_solve_##sfx(const T* lu, int64_t stride_lu, const int32_t* piv, int64_t stride_piv, \
                         const T* b, int64_t stride_b, T* x, int64_t batch, int32_t n,           \
                         int32_t flags, lxb_stream_t stream);                                    \
  int lxb_lu_factor_solve_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b,      \
                                T* x, T* lu, int32_t* piv, int64_t batch, int32_t n,             \
                                lxb_stream_t stream);
LXB_DECL_LU(f32, float)
LXB_DECL_LU(f64, double)

/* ------------------------------------------------------------ Cholesky --
 * lineax/_solver/cholesky.py:43-78.  factor = upper U with (+-A) = U^T U.
 */
#define LXB_DECL_CHOL(sfx, T)                                                                   \
  int lxb_cholesky_factor_##sfx(const T* A, int64_t stride_A, T* factor, int64_t batch,         \
                                int32_t n, int32_t flags, lxb_stream_t stream);                 \
  int lxb_cholesky_solve_##sfx(const T* factor, int64_t stride_f, const T* b, int64_t stride_b, \
                               T* x, int64_t batch, int32_t n, int32_t flags,                   \
                               lxb_stream_t stream);
LXB_DECL_CHOL(f32, float)
LXB_DECL_CHOL(f64, double)

/* ------------------------------------------------------------------ QR --
 * lineax/_solver/qr.py:55-94.  `a`[batch,rows,cols] (rows >= cols) is geqrf's
 * output (R above the diagonal, Householder vectors below), taus[batch,cols].
 * factor: A[batch,m,n]; when n > m the routine factors A^T (qr.py:59-61), i.e.
 *         rows = max(m,n), cols = min(m,n).
 * solve : LXB_TRANS clear -> least squares  x[cols] = R^-1 (Q^T b[rows])[:cols]
 *         LXB_TRANS set   -> minimum norm   x[rows] = Q [R^-T b[cols]; 0]
 */
#define LXB_DECL_QR(sfx, T)                                                                       \
  int lxb_qr_factor_##sfx(const T* A, int64_t stride_A, T* a, T* taus, int64_t batch, int32_t m,  \
                          int32_t n, void* workspace, size_t workspace_bytes,                     \
                          lxb_stream_t stream);                                                   \
  size_t lxb_qr_factor_workspace_##sfx(int64_t batch, int32_t m, int32_t n);                      \
  int lxb_qr_solve_##sfx(const T* a, int64_t stride_a, const T* taus, int64_t stride_t,           \
                         const T* b, int64_t stride_b, T* x, int64_t batch, int32_t rows,         \
                         int32_t cols, int32_t flags, void* workspace, size_t workspace_bytes,    \
                         lxb_stream_t stream);                                                    \
  size_t lxb_qr_solve_workspace_##sfx(int64_t batch, int32_t rows, int32_t cols);
LXB_DECL_QR(f32, float)
LXB_DECL_QR(f64, double)

/* ---------------------------------------------------------- Tridiagonal --
 * lineax/_solver/tridiagonal.py:54-72.  d[batch,n], dl/du[batch,n-1]
 * (dl = sub-diagonal, du = super-diagonal; the reference pads them itself at
 * tridiagonal.py:65-67), b/x[batch,n].  Partial pivoting like LAPACK gtsv.
 * workspace: lxb_tridiagonal_workspace_*() bytes of device scratch (L2-resident slab).
 */
#define LXB_DECL_TRIDIAG(sfx, T)                                                                 \
  int lxb_tridiagonal_solve_##sfx(const T* d, const T* dl, const T* du, int64_t stride_diag,     \
                                  const T* b, int64_t stride_b, T* x, int64_t batch, int32_t n,  \
                                  void* workspace, size_t workspace_bytes, lxb_stream_t stream); \
  size_t lxb_tridiagonal_workspace_##sfx(int64_t batch, int32_t n);
LXB_DECL_TRIDIAG(f32, float)
LXB_DECL_TRIDIAG(f64, double)

/* -------------------------------------------- Diagonal / Triangular -----
 * lineax/_solver/diagonal.py:66-83, triangular.py:68-85 (dispatch targets of
 * AutoLinearSolver, _solve.py:555-600).  rcond < 0 selects the well-posed path.
 */
#define LXB_DECL_DIAGTRI(sfx, T)                                                                 \
  int lxb_diagonal_solve_##sfx(const T* diag, int64_t stride_d, const T* b, int64_t stride_b,    \
                               T* x, int64_t batch, int32_t n, T rcond, lxb_stream_t stream);    \
  int lxb_triangular_solve_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b,     \
                                 T* x, int64_t batch, int32_t n, int32_t flags,                  \
                                 lxb_stream_t stream);
LXB_DECL_DIAGTRI(f32, float)
LXB_DECL_DIAGTRI(f64, double)

/* ---------------------------------------------------------- Krylov ------
 * One persistent fused kernel per batch of systems (matvec + dots + axpys +
 * convergence/breakdown tests, no host round trips).
 *   A[batch,n,n] row-major (lsmr: A[batch,m,n]); b[batch,n]; x[batch,n].
 *   Minv: optional dense preconditioner [batch,n,n] (options["preconditioner"],
 *         _solver/misc.py:30-60) or NULL for the identity.
 *   x: output; with LXB_HAS_Y0 also the initial guess on entry.
 *   result[batch] int32 RESULTS code; num_steps[batch] int32.
 *   max_steps: already resolved by the caller (10*n when None, cg.py:124-127).
 *   workspace: lxb_<solver>_workspace_<sfx>() bytes, only needed for systems
 *              too large to keep their vectors on chip (may be NULL/0 otherwise).
 */
#define LXB_DECL_KRYLOV(sfx, T)                                                                    \
  int lxb_cg_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b, const T* Minv,      \
                   int64_t stride_M, T* x, int32_t* result, int32_t* num_steps, int64_t batch,     \
                   int32_t n, T rtol, T atol, int32_t max_steps, int32_t stabilise_every,          \
                   int32_t flags, void* workspace, size_t workspace_bytes, lxb_stream_t stream);   \
  size_t lxb_cg_workspace_##sfx(int64_t batch, int32_t n);                                         \
  int lxb_bicgstab_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b,               \
                         const T* Minv, int64_t stride_M, T* x, int32_t* result,                   \
                         int32_t* num_steps, int64_t batch, int32_t n, T rtol, T atol,             \
                         int32_t max_steps, int32_t flags, void* workspace,                        \
                         size_t workspace_bytes, lxb_stream_t stream);                             \
  size_t lxb_bicgstab_workspace_##sfx(int64_t batch, int32_t n);                                   \
  int lxb_gmres_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b, const T* Minv,   \
                      int64_t stride_M, T* x, int32_t* result, int32_t* num_steps, int64_t batch,  \
                      int32_t n, T rtol, T atol, int32_t max_steps, int32_t restart,               \
                      int32_t stagnation_iters, int32_t flags, void* workspace,                    \
                      size_t workspace_bytes, lxb_stream_t stream);                                \
  size_t lxb_gmres_workspace_##sfx(int64_t batch, int32_t n, int32_t restart);                     \
  /* stats[batch,8]: istop, norm_r, norm_Ar, norm_A, cond_A, norm_x, 0, 0 (lsmr.py:334-342) */     \
  int lxb_lsmr_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b, T* x,             \
                     int32_t* result, int32_t* num_steps, T* stats, int64_t batch, int32_t m,      \
                     int32_t n, T rtol, T atol, T conlim, int64_t max_steps, int32_t flags,        \
                     void* workspace, size_t workspace_bytes, lxb_stream_t stream);                \
  size_t lxb_lsmr_workspace_##sfx(int64_t batch, int32_t m, int32_t n);
LXB_DECL_KRYLOV(f32, float)
LXB_DECL_KRYLOV(f64, double)

/* ------------------------------------------------ multi-RHS solves -------
 * lineax/_solve.py:732-740 (`state=` reuse), 809-871 (`lx.invert`), and vmap(in_axes=(None, 0)) over
 * `linear_solve`: `nrhs` vectors b[batch, nrhs, n] against ONE factorisation per system; x[batch, nrhs, n].
 * Results are bit-identical to nrhs calls of the single-vector entry points.
 */
#define LXB_DECL_MULTI(sfx, T)                                                                       \
  int lxb_lu_solve_multi_##sfx(const T* lu, int64_t stride_lu, const int32_t* piv, int64_t stride_piv, \
                               const T* b, T* x, int64_t batch, int32_t n, int32_t nrhs, int32_t flags, \
                               lxb_stream_t stream);                                                 \
  int lxb_cholesky_solve_multi_##sfx(const T* factor, int64_t stride_f, const T* b, T* x,            \
                                     int64_t batch, int32_t n, int32_t nrhs, int32_t flags,          \
                                     lxb_stream_t stream);                                           \
  int lxb_triangular_solve_multi_##sfx(const T* A, int64_t stride_A, const T* b, T* x, int64_t batch, \
                                       int32_t n, int32_t nrhs, int32_t flags, lxb_stream_t stream);
LXB_DECL_MULTI(f32, float)
LXB_DECL_MULTI(f64, double)

/* ------------------------------------------------ operator application --
 * MatrixLinearOperator.mv (lineax/_operator.py:265-269; LXB_TRANS -> A^T x),
 * DiagonalLinearOperator.mv (507-511), TridiagonalLinearOperator.mv (861-866;
 * dl/du hold n-1 entries per system, stride_diag counts the diagonal's n) and the
 * vector kernels of lineax/_norm.py:27-139: out[batch,3] = {two_norm(x), max_norm(x),
 * dot(x, y) (0 when y is NULL)}.
 */
#define LXB_DECL_VEC(sfx, T)                                                                      \
  int lxb_matvec_##sfx(const T* A, int64_t stride_A, const T* x, int64_t stride_x, T* y,          \
                       int64_t batch, int32_t m, int32_t n, int32_t flags, lxb_stream_t stream);  \
  int lxb_diag_mv_##sfx(const T* d, int64_t stride_d, const T* x, int64_t stride_x, T* y,         \
                        int64_t batch, int32_t n, lxb_stream_t stream);                           \
  int lxb_tridiag_mv_##sfx(const T* d, const T* dl, const T* du, int64_t stride_diag,             \
                           const T* x, int64_t stride_x, T* y, int64_t batch, int32_t n,          \
                           lxb_stream_t stream);                                                  \
  int lxb_norms_##sfx(const T* x, int64_t stride_x, const T* y, int64_t stride_y, T* out,         \
                      int64_t batch, int64_t n, lxb_stream_t stream);
LXB_DECL_VEC(f32, float)
LXB_DECL_VEC(f64, double)

/* Gram matrix of the normal equations, lineax/_solver/normal.py:111-117 (`conj(op.T) @ op` / `op @ conj(op.T)`):
 * G[batch,n,n] = A^T A (flags = 0) or G[batch,m,m] = A A^T (LXB_TRANS) for A[batch,m,n] row-major. */
#define LXB_DECL_GRAM(sfx, T)                                                                    \
  int lxb_gram_##sfx(const T* A, int64_t stride_A, T* G, int64_t batch, int32_t m, int32_t n,    \
                     int32_t flags, lxb_stream_t stream);
LXB_DECL_GRAM(f32, float)
LXB_DECL_GRAM(f64, double)

/* ------------------------------------------------ multi-GPU, row-sharded --
 * Restarted GMRES on ONE large system partitioned by rows over the GPUs of an NVLink box
 * (one process per GPU).  Rank r owns rows [row_offset, row_offset + n_local) of A
 * (A_local[n_local, n] row-major) and the same slice of b / x.  The per-step all-gather of the
 * Krylov vector and all-reduce of the Gram-Schmidt scalars are fused into the persistent kernel
 * over peer memory: peer_buffers is a DEVICE array of `world` pointers to every rank's symmetric
 * buffer (lxb_gmres_rowsharded_symm_bytes_* bytes each, zero-initialised once, e.g. from
 * torch.distributed._symmetric_memory).  All ranks must call with identical scalar arguments;
 * result / num_steps (1 element) are identical on every rank.
 */
#define LXB_DECL_GMRES_DIST(sfx, T)                                                                 \
  int lxb_gmres_rowsharded_##sfx(const T* A_local, const T* b_local, T* x_local, int32_t* result,  \
                                 int32_t* num_steps, int32_t n, int32_t n_local,                   \
                                 int32_t row_offset, T rtol, T atol, int32_t max_steps,            \
                                 int32_t restart, int32_t stagnation_iters, int32_t flags,         \
                                 void* workspace, size_t workspace_bytes,                          \
                                 void* const* peer_buffers, int32_t world, int32_t rank,           \
                                 lxb_stream_t stream);                                             \
  size_t lxb_gmres_rowsharded_workspace_##sfx(int32_t n_local, int32_t restart);                   \
  size_t lxb_gmres_rowsharded_symm_bytes_##sfx(int32_t n);
LXB_DECL_GMRES_DIST(f32, float)
LXB_DECL_GMRES_DIST(f64, double)

/* CG (lineax/_solver/cg.py:114-227, no preconditioner) on ONE SPD system partitioned by rows like the
 * row-sharded GMRES above: same arguments, the symmetric buffer has lxb_gmres_rowsharded_symm_bytes_* bytes.
 * Three fused exchange rounds per iteration (all-gather of p, <Ap, p>, {<r, r>, the two max-norms}).
 */
#define LXB_DECL_CG_DIST(sfx, T)                                                                    \
  int lxb_cg_rowsharded_##sfx(const T* A_local, const T* b_local, T* x_local, int32_t* result,      \
                              int32_t* num_steps, int32_t n, int32_t n_local, int32_t row_offset,   \
                              T rtol, T atol, int32_t max_steps, int32_t stabilise_every,           \
                              int32_t flags, void* workspace, size_t workspace_bytes,               \
                              void* const* peer_buffers, int32_t world, int32_t rank,               \
                              lxb_stream_t stream);                                                 \
  size_t lxb_cg_rowsharded_workspace_##sfx(int32_t n_local);                                        \
  /* BiCGStab (bicgstab.py:78-205), same conventions and the same workspace / symmetric-buffer sizes */ \
  int lxb_bicgstab_rowsharded_##sfx(const T* A_local, const T* b_local, T* x_local, int32_t* result, \
                                    int32_t* num_steps, int32_t n, int32_t n_local,                  \
                                    int32_t row_offset, T rtol, T atol, int32_t max_steps,           \
                                    int32_t flags, void* workspace, size_t workspace_bytes,          \
                                    void* const* peer_buffers, int32_t world, int32_t rank,          \
                                    lxb_stream_t stream);
LXB_DECL_CG_DIST(f32, float)
LXB_DECL_CG_DIST(f64, double)

/* LSMR (lineax/_solver/lsmr.py:94-409) on ONE tall system partitioned by ROWS over the GPUs of an
 * NVLink box: rank r owns m_local contiguous rows of A (A_local[m_local, n] row-major, 16-byte
 * aligned) and the same slice of b; the solution x[n] and the statistics are replicated (every
 * rank passes its own full-length x; with LXB_HAS_Y0 it holds the same y0 on every rank).  Each
 * iteration reads the local rows once and does one in-kernel exchange over peer memory (all-reduce
 * of the n partial column sums of A^T u and of ||u||^2).  Requires n to be a multiple of 16 bytes
 * worth of elements and n <= 8192 (f32) / 4096 (f64); otherwise LXB_E_UNSUPPORTED.
 * peer_buffers: as for GMRES, lxb_lsmr_rowsharded_symm_bytes_* bytes per rank.
 */
#define LXB_DECL_LSMR_DIST(sfx, T)                                                                  \
  int lxb_lsmr_rowsharded_##sfx(const T* A_local, const T* b_local, T* x, int32_t* result,         \
                                int32_t* num_steps, T* stats, int32_t m, int32_t m_local,          \
                                int32_t n, T rtol, T atol, T conlim, int64_t max_steps,            \
                                int32_t flags, void* workspace, size_t workspace_bytes,            \
                                void* const* peer_buffers, int32_t world, int32_t rank,            \
                                lxb_stream_t stream);                                              \
  size_t lxb_lsmr_rowsharded_workspace_##sfx(int32_t m_local, int32_t n);                          \
  size_t lxb_lsmr_rowsharded_symm_bytes_##sfx(int32_t n, int32_t world);
LXB_DECL_LSMR_DIST(f32, float)
LXB_DECL_LSMR_DIST(f64, double)

/* ------------------------------------------------------ post-processing --
 * lineax/_solve.py:104-123, per system:
 *   successful & any(!isfinite(x)) -> singular;  singular & any(!isfinite(b)) -> nonfinite_input.
 * result[] is updated in place.
 */
#define LXB_DECL_POST(sfx, T)                                                                     \
  int lxb_postprocess_##sfx(const T* x, int64_t stride_x, int32_t nx, const T* b,                 \
                            int64_t stride_b, int32_t nb, int32_t* result, int64_t batch,         \
                            lxb_stream_t stream);
LXB_DECL_POST(f32, float)
LXB_DECL_POST(f64, double)

/* ------------------------------------------------- host-buffer entries --
 * Same computation with HOST pointers: the library stages the batch through
 * device scratch in chunks on internal streams (H2D copy / kernel / D2H copy
 * overlapped).  `device_scratch` (lxb_host_scratch_bytes) is caller-owned
 * device memory.  These are what `bench.py`'s e2e leg times.
 */
int lxb_lu_factor_solve_f32_host(const float* A_host, const float* b_host, float* x_host,
                                 int64_t batch, int32_t n, void* device_scratch,
                                 size_t scratch_bytes, lxb_stream_t stream);
size_t lxb_host_scratch_bytes(int64_t batch, int32_t n, int32_t elem_bytes);

/* fp32 FMA throughput probe: launches `iters` x 64 dependent-chain FMAs per thread on every SM
 * (`packed` != 0: fma.rn.f32x2) and returns the flop count of the launch in *flops (host pointer).
 * bench.py times it with CUDA events to MEASURE the FP32 roofline denominator of the factorisations
 * (BASELINE.md section 2 asks for a measured, not nominal, peak).  `out`: >= 1 float of device memory. */
int lxb_fp32_fma_probe(float* out, int32_t iters, int32_t packed, double* flops, lxb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LINEAX_B200_H_ */

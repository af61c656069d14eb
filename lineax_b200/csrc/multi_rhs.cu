// Multi-right-hand-side solves with ONE factorisation: the `state=` reuse / `lx.invert` / vmap(in_axes=
// (None, 0)) case of lineax/_solve.py:732-740, 809-871 -- many vectors against the same (lu, piv),
// Cholesky factor or triangular operator.
//
// One THREAD per right-hand side, a tile of TR right-hand sides per CTA held in shared memory as
// ys[n][TR + 1]; the factor is read from global memory with warp-uniform (broadcast) loads, so it crosses
// L2 -> SM once per CTA instead of once per vector.  Every element sees exactly the operation sequence of
// the single-vector kernels (lu.cu / direct.cu: y_i = fma(-m_ik, x_k, y_i) with k ascending in a forward
// and descending in a backward substitution, x_k = y_k * (1 / u_kk) for LU, y_k / t_kk otherwise), so the
// results are bit-identical to solving the vectors one by one.
// (The north star's tcgen05 blocked TRSM for this case is not built; see DESIGN.md.)
#include "common.cuh"

namespace lxb {

struct TriPhase {
  int lower;  // stored triangle
  int trans;  // solve with its transpose
  int unit;   // unit diagonal
  int recip;  // x_k = y_k * (1 / t_kk) (LU, getrs order) instead of y_k / t_kk
};
struct MultiPlan {
  int nphases;
  TriPhase ph[2];
  int perm;    // 0 none, 1 apply the getrf row swaps first (forward order), 2 undo them last (reverse order)
  int negate;  // Cholesky of a negative definite operator: x = -x
};

template <typename T>
__device__ __forceinline__ void multi_phase(const T* __restrict__ M, int n, T* ys, int ld, const TriPhase ph) {
  const bool forward = (ph.lower != 0) != (ph.trans != 0);  // effective matrix is lower triangular
  T* y = ys + threadIdx.x;
  if (!ph.trans) {
    // row-oriented: row i of the stored triangle is contiguous
    for (int s = 0; s < n; ++s) {
      const int i = forward ? s : n - 1 - s;
      const T* row = M + (size_t)i * n;
      T acc = y[(size_t)i * ld];
      if (forward) {
        for (int k = 0; k < i; ++k) acc = fma_(-row[k], y[(size_t)k * ld], acc);
      } else {
        for (int k = n - 1; k > i; --k) acc = fma_(-row[k], y[(size_t)k * ld], acc);
      }
      if (!ph.unit) acc = ph.recip ? acc * (T(1) / row[i]) : acc / row[i];
      y[(size_t)i * ld] = acc;
    }
  } else {
    // transposed: row k of the stored triangle holds column k of the effective matrix (axpy form)
    for (int s = 0; s < n; ++s) {
      const int k = forward ? s : n - 1 - s;
      const T* row = M + (size_t)k * n;
      T xk = y[(size_t)k * ld];
      if (!ph.unit) {
        xk = ph.recip ? xk * (T(1) / row[k]) : xk / row[k];
        y[(size_t)k * ld] = xk;
      }
      if (forward) {
        for (int i = k + 1; i < n; ++i) y[(size_t)i * ld] = fma_(-row[i], xk, y[(size_t)i * ld]);
      } else {
        for (int i = 0; i < k; ++i) y[(size_t)i * ld] = fma_(-row[i], xk, y[(size_t)i * ld]);
      }
    }
  }
}

template <typename T>
__global__ void solve_multi_kernel(const T* __restrict__ Mat, int64_t sM, const int32_t* __restrict__ Piv,
                                   int64_t sP, const T* __restrict__ B, T* __restrict__ X, int64_t batch,
                                   int n, int nrhs, MultiPlan plan) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* ys = reinterpret_cast<T*>(smem_raw);
  const int TR = blockDim.x, ld = TR + 1;
  const int tiles = (nrhs + TR - 1) / TR;
  for (int64_t w = blockIdx.x; w < batch * tiles; w += gridDim.x) {
    const int64_t sys = w / tiles;
    const int r0 = (int)(w % tiles) * TR;
    const int cnt = nrhs - r0 < TR ? nrhs - r0 : TR;
    const T* M = Mat + sys * sM;
    const T* b = B + (sys * nrhs + r0) * (int64_t)n;
    T* x = X + (sys * nrhs + r0) * (int64_t)n;
    for (int idx = threadIdx.x; idx < cnt * n; idx += TR) {
      const int r = idx / n, i = idx % n;
      ys[(size_t)i * ld + r] = b[idx];
    }
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      T* y = ys + threadIdx.x;
      if (plan.perm == 1) {
        const int32_t* piv = Piv + sys * sP;
        for (int k = 0; k < n; ++k) {
          const int p = piv[k];
          const T t = y[(size_t)k * ld];
          y[(size_t)k * ld] = y[(size_t)p * ld];
          y[(size_t)p * ld] = t;
        }
      }
      for (int q = 0; q < plan.nphases; ++q) multi_phase<T>(M, n, ys, ld, plan.ph[q]);
      if (plan.perm == 2) {
        const int32_t* piv = Piv + sys * sP;
        for (int k = n - 1; k >= 0; --k) {
          const int p = piv[k];
          const T t = y[(size_t)k * ld];
          y[(size_t)k * ld] = y[(size_t)p * ld];
          y[(size_t)p * ld] = t;
        }
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < cnt * n; idx += TR) {
      const int r = idx / n, i = idx % n;
      const T v = ys[(size_t)i * ld + r];
      x[idx] = plan.negate ? -v : v;
    }
    __syncthreads();
  }
}

template <typename T>
int solve_multi(const T* M, int64_t sM, const int32_t* piv, int64_t sP, const T* b, T* x, int64_t batch, int n,
                int nrhs, const MultiPlan& plan, cudaStream_t st) {
  if (batch < 0 || n < 0 || nrhs < 0 || !M || !b || !x || (plan.perm && !piv)) return LXB_E_BADARG;
  if (batch == 0 || n == 0 || nrhs == 0) return 0;
  int TR = 128;
  while (TR > 32 && ((size_t)n * (TR + 1) * sizeof(T) > 200 * 1024 || TR / 2 >= nrhs)) TR /= 2;
  const size_t smem = (size_t)n * (TR + 1) * sizeof(T);
  if (smem > 220 * 1024) return LXB_E_UNSUPPORTED;
  auto kern = solve_multi_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TR, smem));
  if (occ < 1) occ = 1;
  const int64_t work = batch * ((nrhs + TR - 1) / TR);
  const int64_t cap = (int64_t)kNumSMs * occ;
  kern<<<(unsigned)(work < cap ? work : cap), TR, smem, st>>>(M, sM, piv, sP, b, x, batch, n, nrhs, plan);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_MULTI(sfx, T)                                                                             \
  extern "C" int lxb_lu_solve_multi_##sfx(const T* lu, int64_t stride_lu, const int32_t* piv,             \
                                          int64_t stride_piv, const T* b, T* x, int64_t batch, int32_t n, \
                                          int32_t nrhs, int32_t flags, lxb_stream_t stream) {             \
    lxb::MultiPlan pl{};                                                                                  \
    pl.nphases = 2;                                                                                       \
    if (!(flags & LXB_TRANS)) {                                                                           \
      pl.perm = 1;                                                                                        \
      pl.ph[0] = {1, 0, 1, 0};                                                                            \
      pl.ph[1] = {0, 0, 0, 1};                                                                            \
    } else {                                                                                              \
      pl.perm = 2;                                                                                        \
      pl.ph[0] = {0, 1, 0, 1};                                                                            \
      pl.ph[1] = {1, 1, 1, 0};                                                                            \
    }                                                                                                     \
    return lxb::solve_multi<T>(lu, stride_lu, piv, stride_piv, b, x, batch, n, nrhs, pl,                  \
                               (cudaStream_t)stream);                                                     \
  }                                                                                                       \
  extern "C" int lxb_cholesky_solve_multi_##sfx(const T* factor, int64_t stride_f, const T* b, T* x,      \
                                                int64_t batch, int32_t n, int32_t nrhs, int32_t flags,    \
                                                lxb_stream_t stream) {                                    \
    lxb::MultiPlan pl{};                                                                                  \
    pl.nphases = 2;                                                                                       \
    pl.ph[0] = {0, 1, 0, 0};                                                                              \
    pl.ph[1] = {0, 0, 0, 0};                                                                              \
    pl.negate = (flags & LXB_NSD) ? 1 : 0;                                                                \
    return lxb::solve_multi<T>(factor, stride_f, nullptr, 0, b, x, batch, n, nrhs, pl,                    \
                               (cudaStream_t)stream);                                                     \
  }                                                                                                       \
  extern "C" int lxb_triangular_solve_multi_##sfx(const T* A, int64_t stride_A, const T* b, T* x,         \
                                                  int64_t batch, int32_t n, int32_t nrhs, int32_t flags,  \
                                                  lxb_stream_t stream) {                                  \
    lxb::MultiPlan pl{};                                                                                  \
    pl.nphases = 1;                                                                                       \
    pl.ph[0] = {(flags & LXB_LOWER) ? 1 : 0, (flags & LXB_TRANS) ? 1 : 0, (flags & LXB_UNIT_DIAG) ? 1 : 0, 0}; \
    return lxb::solve_multi<T>(A, stride_A, nullptr, 0, b, x, batch, n, nrhs, pl, (cudaStream_t)stream);  \
  }
LXB_DEF_MULTI(f32, float)
LXB_DEF_MULTI(f64, double)

// Dense direct solvers other than LU, one CTA per system (matrix resident in shared memory
// when it fits, otherwise worked on in place in the caller's output buffer):
//   Cholesky   lineax/_solver/cholesky.py:43-78  (potrf upper + potrs)
//   QR         lineax/_solver/qr.py:55-94        (geqrf + ormqr + trtrs, tall and wide)
//   Triangular lineax/_solver/triangular.py:68-85
//   Diagonal   lineax/_solver/diagonal.py:66-83
#include "common.cuh"

namespace lxb {

constexpr int kDirectThreads = 256;
constexpr size_t kMaxSmemD = 227 * 1024;

// ------------------------------------------------------------- triangular helpers ----
// Solve op(T) x = y in place in y (shared memory).  M row-major, leading dimension ld.
// upper/lower refer to the STORED triangle; trans solves with its transpose.
template <typename T>
__device__ void block_tri_solve(const T* M, int ld, int n, T* y, bool lower, bool trans,
                                bool unit) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool forward = lower != trans;  // effective matrix is lower triangular
  for (int s = 0; s < n; ++s) {
    const int k = forward ? s : n - 1 - s;
    if (!unit) {
      if (tid == 0) y[k] = y[k] / M[(size_t)k * ld + k];
      __syncthreads();
    }
    const T xk = y[k];
    if (forward) {
      for (int i = k + 1 + tid; i < n; i += nt) {
        const T m = trans ? M[(size_t)k * ld + i] : M[(size_t)i * ld + k];
        y[i] = fma_(-m, xk, y[i]);
      }
    } else {
      for (int i = tid; i < k; i += nt) {
        const T m = trans ? M[(size_t)k * ld + i] : M[(size_t)i * ld + k];
        y[i] = fma_(-m, xk, y[i]);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------- Cholesky ----
// Right-looking upper Cholesky: A = U^T U reading the upper triangle of A.
// Non-positive / NaN pivot -> the whole factor becomes NaN (what XLA's potrf returns).
template <typename T>
__global__ void __launch_bounds__(kDirectThreads)
    cholesky_factor_kernel(const T* __restrict__ A, int64_t sA, T* __restrict__ F, int64_t batch,
                           int n, int nsd, int in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  __shared__ int failed;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ld = in_smem ? n + 1 : n;
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    const T* src = A + sys * sA;
    T* out = F + sys * (int64_t)n * n;
    T* M = in_smem ? sm : out;
    const T sign = nsd ? T(-1) : T(1);
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx / n, j = idx % n;
      M[(size_t)i * ld + j] = j >= i ? sign * src[idx] : T(0);
    }
    if (tid == 0) failed = 0;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      const T ajj = M[(size_t)j * ld + j];
      if (!(ajj > T(0))) {
        if (tid == 0) failed = 1;
        break;  // uniform: every thread reads the same ajj
      }
      const T ujj = sqrt_(ajj);
      __syncthreads();
      for (int k = j + tid; k < n; k += nt)
        M[(size_t)j * ld + k] = k == j ? ujj : M[(size_t)j * ld + k] / ujj;
      __syncthreads();
      const int rem = n - j - 1;
      for (int idx = tid; idx < rem * rem; idx += nt) {
        const int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
        if (k >= i)
          M[(size_t)i * ld + k] = fma_(-M[(size_t)j * ld + i], M[(size_t)j * ld + k], M[(size_t)i * ld + k]);
      }
      __syncthreads();
    }
    __syncthreads();
    const bool bad = failed != 0;
    if (in_smem || bad) {
      for (int idx = tid; idx < n * n; idx += nt)
        out[idx] = bad ? Num<T>::nan() : M[(size_t)(idx / n) * ld + idx % n];
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(kDirectThreads)
    cholesky_solve_kernel(const T* __restrict__ F, int64_t sF, const T* __restrict__ B, int64_t sB,
                          T* __restrict__ X, int64_t batch, int n, int nsd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* y = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    const T* M = F + sys * sF;
    for (int i = tid; i < n; i += nt) y[i] = B[sys * sB + i];
    __syncthreads();
    block_tri_solve<T>(M, n, n, y, /*lower=*/false, /*trans=*/true, false);   // U^T z = b
    block_tri_solve<T>(M, n, n, y, /*lower=*/false, /*trans=*/false, false);  // U x = z
    for (int i = tid; i < n; i += nt) X[sys * n + i] = nsd ? -y[i] : y[i];
    __syncthreads();
  }
}

// -------------------------------------------------------------------- Triangular ----
template <typename T>
__global__ void __launch_bounds__(kDirectThreads)
    triangular_solve_kernel(const T* __restrict__ A, int64_t sA, const T* __restrict__ B,
                            int64_t sB, T* __restrict__ X, int64_t batch, int n, int flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* y = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    for (int i = tid; i < n; i += nt) y[i] = B[sys * sB + i];
    __syncthreads();
    block_tri_solve<T>(A + sys * sA, n, n, y, (flags & LXB_LOWER) != 0, (flags & LXB_TRANS) != 0,
                       (flags & LXB_UNIT_DIAG) != 0);
    for (int i = tid; i < n; i += nt) X[sys * n + i] = y[i];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------- Diagonal ----
// diagonal.py:66-83: rcond < 0 -> well-posed (x = b / d); else x = b / where(|d| > rcond*max|d|, d, inf)
template <typename T>
__global__ void __launch_bounds__(kDirectThreads)
    diagonal_solve_kernel(const T* __restrict__ D, int64_t sD, const T* __restrict__ B, int64_t sB,
                          T* __restrict__ X, int64_t batch, int n, T rcond) {
  __shared__ T red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    const T* d = D + sys * sD;
    T cut = T(-1);
    if (rcond >= T(0)) {
      T mx[1] = {T(0)};
      for (int i = tid; i < n; i += nt) mx[0] = absmax2(mx[0], d[i]);
      block_absmax<T, 1>(mx, red);
      cut = rcond * mx[0];
    }
    for (int i = tid; i < n; i += nt) {
      const T di = d[i];
      const T den = (rcond < T(0) || abs_(di) > cut) ? di : Num<T>::inf();
      X[sys * n + i] = B[sys * sB + i] / den;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------- QR ----
// Unblocked Householder QR (LAPACK geqr2 conventions) of a rows x cols matrix, rows >= cols.
// Column reductions: a warp owns 32 consecutive columns (lane = column), warps split the rows.
template <typename T>
__device__ void block_geqr2(T* M, int ld, int rows, int cols, T* taus, T* red, T* part) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int j = 0; j < cols; ++j) {
    // larfg on column j
    T ssq[1] = {T(0)};
    for (int i = j + 1 + tid; i < rows; i += nt) {
      const T v = M[(size_t)i * ld + j];
      ssq[0] = fma_(v, v, ssq[0]);
    }
    block_sum<T, 1>(ssq, red);
    const T alpha = M[(size_t)j * ld + j];
    T tau = T(0), beta = alpha, scal = T(1);
    if (ssq[0] != T(0)) {
      const T nrm = sqrt_(alpha * alpha + ssq[0]);
      beta = alpha >= T(0) ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scal = T(1) / (alpha - beta);
    }
    __syncthreads();
    if (ssq[0] != T(0))
      for (int i = j + 1 + tid; i < rows; i += nt) M[(size_t)i * ld + j] *= scal;
    if (tid == 0) {
      M[(size_t)j * ld + j] = beta;
      taus[j] = tau;
    }
    __syncthreads();
    if (tau != T(0)) {
      // w_c = a_jc + sum_i v_i a_ic ;  a_:c -= tau w_c [1; v]
      for (int c0 = j + 1; c0 < cols; c0 += 32) {
        const int c = c0 + lane;
        T acc = T(0);
        if (c < cols)
          for (int i = j + 1 + warp; i < rows; i += nw)
            acc = fma_(M[(size_t)i * ld + j], M[(size_t)i * ld + c], acc);
        part[warp * 33 + lane] = acc;
        __syncthreads();
        if (c < cols) {
          T w = M[(size_t)j * ld + c];
          for (int q = 0; q < nw; ++q) w += part[q * 33 + lane];
          const T f = tau * w;
          for (int i = j + 1 + warp; i < rows; i += nw)
            M[(size_t)i * ld + c] = fma_(-f, M[(size_t)i * ld + j], M[(size_t)i * ld + c]);
          if (warp == 0) M[(size_t)j * ld + c] -= f;
        }
        __syncthreads();
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kDirectThreads)
    qr_factor_kernel(const T* __restrict__ A, int64_t sA, T* __restrict__ Aout, T* __restrict__ Taus,
                     int64_t batch, int m, int n, int in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool wide = n > m;  // qr.py:59-61: factor A^T
  const int rows = wide ? n : m, cols = wide ? m : n;
  T* red = sm;            // 32
  T* part = red + 32;     // 8 * 33
  T* staus = part + 8 * 33;  // cols
  T* Ms = staus + ((cols + 3) & ~3);
  const int ld = in_smem ? cols + 1 : cols;
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    const T* src = A + sys * sA;
    T* out = Aout + sys * (int64_t)rows * cols;
    T* M = in_smem ? Ms : out;
    for (int idx = tid; idx < rows * cols; idx += nt) {
      const int i = idx / cols, c = idx % cols;
      M[(size_t)i * ld + c] = wide ? src[(size_t)c * n + i] : src[idx];
    }
    __syncthreads();
    block_geqr2<T>(M, ld, rows, cols, staus, red, part);
    if (in_smem)
      for (int idx = tid; idx < rows * cols; idx += nt) out[idx] = M[(size_t)(idx / cols) * ld + idx % cols];
    for (int j = tid; j < cols; j += nt) Taus[sys * cols + j] = staus[j];
    __syncthreads();
  }
}

// y (rows, shared) <- H_j y for one reflector j stored in column j of `a`.
template <typename T>
__device__ __forceinline__ void apply_reflector(const T* a, int cols, int rows, int j, T tau, T* y,
                                                T* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  T d[1] = {T(0)};
  for (int i = j + 1 + tid; i < rows; i += nt) d[0] = fma_(a[(size_t)i * cols + j], y[i], d[0]);
  block_sum<T, 1>(d, red);
  const T f = tau * (y[j] + d[0]);
  __syncthreads();
  for (int i = j + 1 + tid; i < rows; i += nt) y[i] = fma_(-f, a[(size_t)i * cols + j], y[i]);
  if (tid == 0) y[j] -= f;
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kDirectThreads)
    qr_solve_kernel(const T* __restrict__ Aq, int64_t sAq, const T* __restrict__ Taus, int64_t sT,
                    const T* __restrict__ B, int64_t sB, T* __restrict__ X, int64_t batch, int rows,
                    int cols, int trans, int qt_only) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* red = reinterpret_cast<T*>(smem_raw);  // 32
  T* y = red + 32;                          // rows
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    const T* a = Aq + sys * sAq;
    const T* taus = Taus + sys * sT;
    if (!trans) {
      // least squares: x = R^-1 (Q^T b)[:cols]   (qr.py:89-92)
      for (int i = tid; i < rows; i += nt) y[i] = B[sys * sB + i];
      __syncthreads();
      for (int j = 0; j < cols; ++j) apply_reflector<T>(a, cols, rows, j, taus[j], y, red);
      if (!qt_only) block_tri_solve<T>(a, cols, cols, y, false, false, false);
      for (int i = tid; i < cols; i += nt) X[sys * cols + i] = y[i];
    } else {
      // minimum norm: x = Q [R^-T b; 0]   (qr.py:79-86)
      for (int i = tid; i < rows; i += nt) y[i] = i < cols ? B[sys * sB + i] : T(0);
      __syncthreads();
      block_tri_solve<T>(a, cols, cols, y, false, true, false);
      for (int j = cols - 1; j >= 0; --j) apply_reflector<T>(a, cols, rows, j, taus[j], y, red);
      for (int i = tid; i < rows; i += nt) X[sys * rows + i] = y[i];
    }
    __syncthreads();
  }
}

// large single tall matrix: blocked Householder on all SMs (qr_large.cu)
inline bool qr_use_large(int64_t batch, int rows, int cols) {
  return batch == 1 && rows >= 4096 && cols >= 64 && (int64_t)rows * cols >= (int64_t)1 << 22;
}
template <typename T>
int qr_large_factor(const T* A, T* a, T* taus, int m, int n, void* ws, size_t ws_bytes, cudaStream_t st);
template <typename T>
int qr_large_solve(const T* a, const T* taus, const T* b, T* x, int m, int n, void* ws, size_t ws_bytes,
                   cudaStream_t st, bool qt_only);
template <typename T>
size_t qr_large_ws_bytes(int m, int n);

inline int64_t grid_for(int64_t batch, int occ) {
  const int64_t cap = (int64_t)kNumSMs * (occ < 1 ? 1 : occ);
  return batch < cap ? batch : cap;
}

template <typename K>
int occupancy(K kern, int threads, size_t smem, int* occ) {
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, threads, smem));
  return 0;
}

template <typename T>
int cholesky_factor(const T* A, int64_t sA, T* F, int64_t batch, int n, int flags, cudaStream_t st) {
  if (batch < 0 || n < 0 || !A || !F) return LXB_E_BADARG;
  if (batch == 0 || n == 0) return 0;
  const size_t mat = (size_t)n * (n + 1) * sizeof(T);
  const int in_smem = mat <= kMaxSmemD - 1024;
  const size_t smem = in_smem ? mat : 16;
  int occ = 1, rc = occupancy(cholesky_factor_kernel<T>, kDirectThreads, smem, &occ);
  if (rc) return rc;
  cholesky_factor_kernel<T><<<(unsigned)grid_for(batch, occ), kDirectThreads, smem, st>>>(
      A, sA, F, batch, n, (flags & LXB_NSD) ? 1 : 0, in_smem);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int cholesky_solve(const T* F, int64_t sF, const T* b, int64_t sb, T* x, int64_t batch, int n,
                   int flags, cudaStream_t st) {
  if (batch < 0 || n < 0 || !F || !b || !x) return LXB_E_BADARG;
  if (batch == 0 || n == 0) return 0;
  const size_t smem = (size_t)n * sizeof(T);
  if (smem > kMaxSmemD) return LXB_E_UNSUPPORTED;
  int occ = 1, rc = occupancy(cholesky_solve_kernel<T>, kDirectThreads, smem, &occ);
  if (rc) return rc;
  cholesky_solve_kernel<T><<<(unsigned)grid_for(batch, occ), kDirectThreads, smem, st>>>(
      F, sF, b, sb, x, batch, n, (flags & LXB_NSD) ? 1 : 0);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int triangular_solve(const T* A, int64_t sA, const T* b, int64_t sb, T* x, int64_t batch, int n,
                     int flags, cudaStream_t st) {
  if (batch < 0 || n < 0 || !A || !b || !x) return LXB_E_BADARG;
  if (batch == 0 || n == 0) return 0;
  const size_t smem = (size_t)n * sizeof(T);
  if (smem > kMaxSmemD) return LXB_E_UNSUPPORTED;
  int occ = 1, rc = occupancy(triangular_solve_kernel<T>, kDirectThreads, smem, &occ);
  if (rc) return rc;
  triangular_solve_kernel<T><<<(unsigned)grid_for(batch, occ), kDirectThreads, smem, st>>>(
      A, sA, b, sb, x, batch, n, flags);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int diagonal_solve(const T* d, int64_t sd, const T* b, int64_t sb, T* x, int64_t batch, int n,
                   T rcond, cudaStream_t st) {
  if (batch < 0 || n < 0 || !d || !b || !x) return LXB_E_BADARG;
  if (batch == 0 || n == 0) return 0;
  diagonal_solve_kernel<T><<<(unsigned)grid_for(batch, 8), kDirectThreads, 0, st>>>(d, sd, b, sb, x,
                                                                                    batch, n, rcond);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int qr_factor(const T* A, int64_t sA, T* a, T* taus, int64_t batch, int m, int n, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  if (batch < 0 || m < 0 || n < 0 || !A || !a || !taus) return LXB_E_BADARG;
  if (batch == 0 || m == 0 || n == 0) return 0;
  const int rows = m > n ? m : n, cols = m > n ? n : m;
  if (m >= n && qr_use_large(batch, rows, cols)) return qr_large_factor<T>(A, a, taus, m, n, ws, ws_bytes, st);
  const size_t fixed = (32 + 8 * 33 + ((cols + 3) & ~3)) * sizeof(T);
  if (fixed > kMaxSmemD) return LXB_E_UNSUPPORTED;
  const size_t mat = (size_t)rows * (cols + 1) * sizeof(T);
  const int in_smem = fixed + mat <= kMaxSmemD;
  const size_t smem = fixed + (in_smem ? mat : 0);
  int occ = 1, rc = occupancy(qr_factor_kernel<T>, kDirectThreads, smem, &occ);
  if (rc) return rc;
  qr_factor_kernel<T><<<(unsigned)grid_for(batch, occ), kDirectThreads, smem, st>>>(A, sA, a, taus, batch,
                                                                                   m, n, in_smem);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int qr_solve(const T* a, int64_t sa, const T* taus, int64_t stau, const T* b, int64_t sb, T* x,
             int64_t batch, int rows, int cols, int flags, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (batch < 0 || rows < 0 || cols < 0 || !a || !taus || !b || !x) return LXB_E_BADARG;
  if (batch == 0 || rows == 0 || cols == 0) return 0;
  if (!(flags & LXB_TRANS) && qr_use_large(batch, rows, cols))
    return qr_large_solve<T>(a, taus, b, x, rows, cols, ws, ws_bytes, st, (flags & LXB_QT_ONLY) != 0);
  const size_t smem = (32 + (size_t)rows) * sizeof(T);
  if (smem > kMaxSmemD) return LXB_E_UNSUPPORTED;
  int occ = 1, rc = occupancy(qr_solve_kernel<T>, kDirectThreads, smem, &occ);
  if (rc) return rc;
  qr_solve_kernel<T><<<(unsigned)grid_for(batch, occ), kDirectThreads, smem, st>>>(
      a, sa, taus, stau, b, sb, x, batch, rows, cols, (flags & LXB_TRANS) ? 1 : 0, (flags & LXB_QT_ONLY) ? 1 : 0);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_DIRECT(sfx, T)                                                                     \
  extern "C" int lxb_cholesky_factor_##sfx(const T* A, int64_t stride_A, T* factor, int64_t batch, \
                                           int32_t n, int32_t flags, lxb_stream_t stream) {        \
    return lxb::cholesky_factor<T>(A, stride_A, factor, batch, n, flags, (cudaStream_t)stream);    \
  }                                                                                                \
  extern "C" int lxb_cholesky_solve_##sfx(const T* factor, int64_t stride_f, const T* b,           \
                                          int64_t stride_b, T* x, int64_t batch, int32_t n,        \
                                          int32_t flags, lxb_stream_t stream) {                    \
    return lxb::cholesky_solve<T>(factor, stride_f, b, stride_b, x, batch, n, flags,               \
                                  (cudaStream_t)stream);                                           \
  }                                                                                                \
  extern "C" int lxb_triangular_solve_##sfx(const T* A, int64_t stride_A, const T* b,              \
                                            int64_t stride_b, T* x, int64_t batch, int32_t n,      \
                                            int32_t flags, lxb_stream_t stream) {                  \
    return lxb::triangular_solve<T>(A, stride_A, b, stride_b, x, batch, n, flags,                  \
                                    (cudaStream_t)stream);                                         \
  }                                                                                                \
  extern "C" int lxb_diagonal_solve_##sfx(const T* diag, int64_t stride_d, const T* b,             \
                                          int64_t stride_b, T* x, int64_t batch, int32_t n,        \
                                          T rcond, lxb_stream_t stream) {                          \
    return lxb::diagonal_solve<T>(diag, stride_d, b, stride_b, x, batch, n, rcond,                 \
                                  (cudaStream_t)stream);                                           \
  }                                                                                                \
  extern "C" int lxb_qr_factor_##sfx(const T* A, int64_t stride_A, T* a, T* taus, int64_t batch,   \
                                     int32_t m, int32_t n, void* workspace,                        \
                                     size_t workspace_bytes, lxb_stream_t stream) {                \
    return lxb::qr_factor<T>(A, stride_A, a, taus, batch, m, n, workspace, workspace_bytes,        \
                             (cudaStream_t)stream);                                                \
  }                                                                                                \
  extern "C" size_t lxb_qr_factor_workspace_##sfx(int64_t batch, int32_t m, int32_t n) {           \
    return (m >= n && lxb::qr_use_large(batch, m, n)) ? lxb::qr_large_ws_bytes<T>(m, n) : 0;       \
  }                                                                                                \
  extern "C" int lxb_qr_solve_##sfx(const T* a, int64_t stride_a, const T* taus, int64_t stride_t, \
                                    const T* b, int64_t stride_b, T* x, int64_t batch,             \
                                    int32_t rows, int32_t cols, int32_t flags, void* workspace,    \
                                    size_t workspace_bytes, lxb_stream_t stream) {                 \
    return lxb::qr_solve<T>(a, stride_a, taus, stride_t, b, stride_b, x, batch, rows, cols, flags, \
                            workspace, workspace_bytes, (cudaStream_t)stream);                     \
  }                                                                                                \
  extern "C" size_t lxb_qr_solve_workspace_##sfx(int64_t batch, int32_t rows, int32_t cols) {      \
    return lxb::qr_use_large(batch, rows, cols) ? lxb::qr_large_ws_bytes<T>(rows, cols) : 0;       \
  }
LXB_DEF_DIRECT(f32, float)
LXB_DEF_DIRECT(f64, double)

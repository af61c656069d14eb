// Operator application and vector kernels outside the fused solvers:
//   MatrixLinearOperator.mv  (lineax/_operator.py:265-269)   -> lxb_matvec_*
//   DiagonalLinearOperator.mv (507-511), TridiagonalLinearOperator.mv (861-866)
//   tree_dot / two_norm / max_norm (lineax/_norm.py:27-139)  -> lxb_dot_*, lxb_norms_*
#include "common.cuh"

namespace lxb {

// y[b, i] = sum_j A[b, i, j] x[b, j]: one warp per row, 128-bit streaming loads of A.
template <typename T>
__global__ void __launch_bounds__(256)
    matvec_kernel(const T* __restrict__ A, int64_t sA, const T* __restrict__ x, int64_t sx,
                  T* __restrict__ y, int64_t batch, int m, int n, int vec) {
  using VT = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
  constexpr int V = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = batch * m;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const int64_t sys = r / m;
    const int i = (int)(r % m);
    const T* row = A + sys * sA + (int64_t)i * n;
    const T* xv = x + sys * sx;
    T acc0 = T(0), acc1 = T(0);
    if (vec) {
      const VT* r4 = reinterpret_cast<const VT*>(row);
      const VT* x4 = reinterpret_cast<const VT*>(xv);
      const int nv = n / V;
      int c = lane;
      for (; c + 32 < nv; c += 64) {
        const VT a0 = ldg_stream(r4 + c), a1 = ldg_stream(r4 + c + 32);
        const VT b0 = x4[c], b1 = x4[c + 32];
        const T* pa0 = reinterpret_cast<const T*>(&a0);
        const T* pb0 = reinterpret_cast<const T*>(&b0);
        const T* pa1 = reinterpret_cast<const T*>(&a1);
        const T* pb1 = reinterpret_cast<const T*>(&b1);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          acc0 = fma_(pa0[e], pb0[e], acc0);
          acc1 = fma_(pa1[e], pb1[e], acc1);
        }
      }
      for (; c < nv; c += 32) {
        const VT a0 = ldg_stream(r4 + c);
        const VT b0 = x4[c];
        const T* pa0 = reinterpret_cast<const T*>(&a0);
        const T* pb0 = reinterpret_cast<const T*>(&b0);
#pragma unroll
        for (int e = 0; e < V; ++e) acc0 = fma_(pa0[e], pb0[e], acc0);
      }
    } else {
      for (int c = lane; c < n; c += 32) acc0 = fma_(row[c], xv[c], acc0);
    }
    const T s = warp_sum(acc0 + acc1);
    if (lane == 0) y[sys * m + i] = s;
  }
}

// y[b, j] = sum_i A[b, i, j] x[b, i]  (A^T x): a CTA owns (system, 32-column tile); 8 warps
// split the rows, partial sums are combined through shared memory in a fixed order.
template <typename T>
__global__ void __launch_bounds__(256)
    matvec_t_kernel(const T* __restrict__ A, int64_t sA, const T* __restrict__ x, int64_t sx,
                    T* __restrict__ y, int64_t batch, int m, int n) {
  __shared__ T part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles = (n + 31) / 32;
  const int64_t total = batch * tiles;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int64_t sys = t / tiles;
    const int j = (int)(t % tiles) * 32 + lane;
    const T* Ab = A + sys * sA;
    const T* xv = x + sys * sx;
    T acc = T(0);
    if (j < n)
      for (int i = warp; i < m; i += 8) acc = fma_(Ab[(int64_t)i * n + j], xv[i], acc);
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && j < n) {
      T s = part[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) s += part[w][lane];
      y[sys * n + j] = s;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void diag_mv_kernel(const T* __restrict__ d, int64_t sd, const T* __restrict__ x,
                               int64_t sx, T* __restrict__ y, int64_t batch, int n) {
  const int64_t total = batch * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t sys = idx / n;
    const int i = (int)(idx % n);
    y[idx] = d[sys * sd + i] * x[sys * sx + i];
  }
}

// lineax/_operator.py:861-866: b.at[:-1].add(a).at[1:].add(c)  ->  (d*v + u*v[+1]) + l*v[-1]
template <typename T>
__global__ void tridiag_mv_kernel(const T* __restrict__ d, const T* __restrict__ dl,
                                  const T* __restrict__ du, int64_t sd, const T* __restrict__ x,
                                  int64_t sx, T* __restrict__ y, int64_t batch, int n) {
  const int64_t total = batch * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t sys = idx / n;
    const int i = (int)(idx % n);
    const int64_t od = sys * sd - (sd ? sys : 0);  // off-diagonals have n-1 entries per system
    const T* xv = x + sys * sx;
    T v = d[sys * sd + i] * xv[i];
    if (i + 1 < n) v = v + du[od + i] * xv[i + 1];
    if (i > 0) v = v + dl[od + i - 1] * xv[i - 1];
    y[idx] = v;
  }
}

// out[b, 0] = two_norm(x_b) (size-1 shortcut |x|), out[b, 1] = max_norm(x_b) (NaN-propagating),
// out[b, 2] = dot(x_b, y_b) when y != null.  One CTA per vector.
template <typename T>
__global__ void __launch_bounds__(256)
    norms_kernel(const T* __restrict__ x, int64_t sx, const T* __restrict__ y, int64_t sy,
                 T* __restrict__ out, int64_t batch, int64_t n) {
  __shared__ T red[96];
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    const T* xv = x + sys * sx;
    const T* yv = y ? y + sys * sy : nullptr;
    T s[2] = {T(0), T(0)};
    T mx[1] = {T(0)};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const T v = xv[i];
      s[0] = fma_(v, v, s[0]);
      if (yv) s[1] = fma_(v, yv[i], s[1]);
      mx[0] = absmax2(mx[0], v);
    }
    block_sum<T, 2>(s, red);
    block_absmax<T, 1>(mx, red + 64);
    if (threadIdx.x == 0) {
      out[sys * 3 + 0] = n == 1 ? abs_(xv[0]) : sqrt_(s[0]);
      out[sys * 3 + 1] = mx[0];
      out[sys * 3 + 2] = s[1];
    }
    __syncthreads();
  }
}

template <typename T>
int matvec(const T* A, int64_t sA, const T* x, int64_t sx, T* y, int64_t batch, int m, int n,
           int flags, cudaStream_t st) {
  if (batch < 0 || m < 0 || n < 0 || !A || !x || !y) return LXB_E_BADARG;
  if (batch == 0 || m == 0) return 0;
  if (flags & LXB_TRANS) {
    const int64_t tiles = batch * ((n + 31) / 32);
    int64_t blocks = tiles < (int64_t)kNumSMs * 8 ? tiles : (int64_t)kNumSMs * 8;
    if (blocks == 0) return 0;
    matvec_t_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(A, sA, x, sx, y, batch, m, n);
  } else {
    constexpr int V = 16 / sizeof(T);
    const int vec = (n % V == 0) && aligned16(A) && aligned16(x) && (sA % V == 0) && (sx % V == 0);
    const int64_t rows = batch * m;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    matvec_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(A, sA, x, sx, y, batch, m, n, vec);
  }
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_VEC(sfx, T)                                                                        \
  extern "C" int lxb_matvec_##sfx(const T* A, int64_t stride_A, const T* x, int64_t stride_x,      \
                                  T* y, int64_t batch, int32_t m, int32_t n, int32_t flags,        \
                                  lxb_stream_t stream) {                                           \
    return lxb::matvec<T>(A, stride_A, x, stride_x, y, batch, m, n, flags, (cudaStream_t)stream);  \
  }                                                                                                \
  extern "C" int lxb_diag_mv_##sfx(const T* d, int64_t stride_d, const T* x, int64_t stride_x,     \
                                   T* y, int64_t batch, int32_t n, lxb_stream_t stream) {          \
    if (batch < 0 || n < 0 || !d || !x || !y) return LXB_E_BADARG;                                 \
    if (batch * n == 0) return 0;                                                                  \
    int64_t blocks = (batch * n + 255) / 256;                                                      \
    if (blocks > lxb::kNumSMs * 8) blocks = lxb::kNumSMs * 8;                                      \
    lxb::diag_mv_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(                    \
        d, stride_d, x, stride_x, y, batch, n);                                                    \
    LXB_CUDA_CHECK_LAUNCH();                                                                       \
    return 0;                                                                                      \
  }                                                                                                \
  extern "C" int lxb_tridiag_mv_##sfx(const T* d, const T* dl, const T* du, int64_t stride_diag,   \
                                      const T* x, int64_t stride_x, T* y, int64_t batch,           \
                                      int32_t n, lxb_stream_t stream) {                            \
    if (batch < 0 || n < 0 || !d || !x || !y || (n > 1 && (!dl || !du))) return LXB_E_BADARG;      \
    if (batch * n == 0) return 0;                                                                  \
    int64_t blocks = (batch * n + 255) / 256;                                                      \
    if (blocks > lxb::kNumSMs * 8) blocks = lxb::kNumSMs * 8;                                      \
    lxb::tridiag_mv_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(                 \
        d, dl, du, stride_diag, x, stride_x, y, batch, n);                                         \
    LXB_CUDA_CHECK_LAUNCH();                                                                       \
    return 0;                                                                                      \
  }                                                                                                \
  extern "C" int lxb_norms_##sfx(const T* x, int64_t stride_x, const T* y, int64_t stride_y,       \
                                 T* out, int64_t batch, int64_t n, lxb_stream_t stream) {          \
    if (batch < 0 || n < 0 || !x || !out) return LXB_E_BADARG;                                     \
    if (batch == 0) return 0;                                                                      \
    int64_t blocks = batch < lxb::kNumSMs * 8 ? batch : lxb::kNumSMs * 8;                          \
    lxb::norms_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(                      \
        x, stride_x, y, stride_y, out, batch, n);                                                  \
    LXB_CUDA_CHECK_LAUNCH();                                                                       \
    return 0;                                                                                      \
  }
LXB_DEF_VEC(f32, float)
LXB_DEF_VEC(f64, double)

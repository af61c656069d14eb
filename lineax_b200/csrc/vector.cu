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

// ---- fp32 FMA throughput probe -------------------------------------------------------------------
// bench.py measures the FLOP roofline denominator of the factorisations on the box it runs on
//  instead of quoting the nominal 148 x 128 x 2 x 1.965 GHz.
namespace lxb {
template <bool PACKED>
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float seed) {
  if constexpr (PACKED) {
    unsigned long long acc[8], m, c;
    asm("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(seed));
    asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(seed * 0.5f));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = seed + (float)(threadIdx.x + j);
      asm("mov.b64 %0, {%1, %1};" : "=l"(acc[j]) : "f"(v));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[j]) : "l"(m), "l"(c));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= acc[j];
    if (s == 0x12345678ull) out[0] = 1.f;  // never true: keeps the chain alive
  } else {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = seed + (float)(threadIdx.x + j);
    const float m = seed, c = seed * 0.5f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[j]) : "f"(m), "f"(c));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j];
    if (s == 1234.5678f) out[0] = s;
  }
}
}  // namespace lxb

extern "C" int lxb_fp32_fma_probe(float* out, int32_t iters, int32_t packed, double* flops,
                                  lxb_stream_t stream) {
  if (out == nullptr || iters < 1) return LXB_E_BADARG;
  const int blocks = lxb::kNumSMs * 8, threads = 256;
  if (packed)
    lxb::fma_probe_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters, 0.999f);
  else
    lxb::fma_probe_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters, 0.999f);
  LXB_CUDA_CHECK_LAUNCH();
  if (flops) *flops = (double)blocks * threads * (double)iters * 64.0 * 2.0 * (packed ? 2.0 : 1.0);
  return 0;
}

// ---- Gram matrices for the normal equations (lineax/_solver/normal.py:111-117) -------------------
// G = A^T A (LXB_TRANS clear, n x n) or A A^T (LXB_TRANS set, m x m) for A[batch, m, n] row-major:
// 64 x 64 output tile per CTA, 16 x 16 threads with a 4 x 4 register tile, 16-deep shared-memory
// panels; only tiles on or above the diagonal are computed and mirrored (G is symmetric).
namespace lxb {
template <typename T>
__global__ void __launch_bounds__(256)
    gram_kernel(const T* __restrict__ A, int64_t sA, T* __restrict__ G, int64_t batch, int m, int n,
                int aat) {
  constexpr int TS = 64, KS = 16;
  __shared__ T Ps[KS][TS + 4], Qs[KS][TS + 4];
  const int g = aat ? m : n;    // order of G
  const int kd = aat ? n : m;   // contraction length
  const int tiles = (g + TS - 1) / TS;
  const int ntri = tiles * (tiles + 1) / 2;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int64_t w = blockIdx.x; w < batch * ntri; w += gridDim.x) {
    const int64_t sys = w / ntri;
    int t = (int)(w % ntri), ti = 0;
    while (t >= tiles - ti) { t -= tiles - ti; ++ti; }
    const int tj = ti + t;  // tile (ti, tj), tj >= ti
    const T* a = A + sys * sA;
    // element (row r of G, contraction index k): aat ? a[r * n + k] : a[k * n + r]
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    for (int k0 = 0; k0 < kd; k0 += KS) {
      for (int e = threadIdx.x; e < KS * TS; e += 256) {
        int kk, rr;
        if (aat) { kk = e % KS; rr = e / KS; } else { rr = e % TS; kk = e / TS; }
        const int k = k0 + kk, ri = ti * TS + rr, rj = tj * TS + rr;
        Ps[kk][rr] = (k < kd && ri < g) ? (aat ? a[(size_t)ri * n + k] : a[(size_t)k * n + ri]) : T(0);
        Qs[kk][rr] = (k < kd && rj < g) ? (aat ? a[(size_t)rj * n + k] : a[(size_t)k * n + rj]) : T(0);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        T p[4], q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = Ps[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = Qs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma_(p[i], q[j], acc[i][j]);
      }
      __syncthreads();
    }
    T* gout = G + sys * (int64_t)g * g;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = ti * TS + ty * 4 + i, c = tj * TS + tx * 4 + j;
        if (r < g && c < g) {
          gout[(size_t)r * g + c] = acc[i][j];
          if (ti != tj) gout[(size_t)c * g + r] = acc[i][j];
        }
      }
  }
}
}  // namespace lxb

#define LXB_DEF_GRAM(sfx, T)                                                                       \
  extern "C" int lxb_gram_##sfx(const T* A, int64_t stride_A, T* G, int64_t batch, int32_t m,      \
                                int32_t n, int32_t flags, lxb_stream_t stream) {                   \
    if (batch < 0 || m < 0 || n < 0 || !A || !G) return LXB_E_BADARG;                              \
    const int aat = (flags & LXB_TRANS) ? 1 : 0;                                                   \
    const int g = aat ? m : n;                                                                     \
    if (batch == 0 || g == 0) return 0;                                                            \
    const int64_t tiles = (g + 63) / 64, work = batch * tiles * (tiles + 1) / 2;                   \
    const int64_t blocks = work < (int64_t)lxb::kNumSMs * 4 ? work : (int64_t)lxb::kNumSMs * 4;    \
    lxb::gram_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, stride_A, G, batch, \
                                                                            m, n, aat);            \
    LXB_CUDA_CHECK_LAUNCH();                                                                       \
    return 0;                                                                                      \
  }
LXB_DEF_GRAM(f32, float)
LXB_DEF_GRAM(f64, double)

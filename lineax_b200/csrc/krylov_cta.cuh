// Building blocks for the "one CTA per system" persistent Krylov kernels.
// The system's vectors live in shared memory for the whole solve; A is either staged into
// shared memory once per solve (A_SMEM) or streamed from global/L2 on every matvec.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace lxb {

constexpr int kKrylovThreads = 256;

template <typename T>
struct KrylovParams {
  const T* A;
  int64_t sA;
  const T* b;
  int64_t sb;
  const T* M;  // optional preconditioner matrix (applied as M @ r), may be null
  int64_t sM;
  T* x;
  int32_t* result;
  int32_t* num_steps;
  T* stats;  // lsmr only
  int64_t batch;
  int m, n;  // rows, cols (square solvers: m == n)
  T rtol, atol, conlim;
  int64_t max_steps;
  int stabilise_every;  // cg: 0 = never (None), 1 = always, k = every k steps
  int restart, stagnation_iters;
  int flags;
  int a_smem;  // stage A in shared memory
  T* ws;       // global workspace (per-system slices), may be null
  int64_t ws_stride;
};

template <typename T>
struct V16K;
template <>
struct V16K<float> {
  using type = float4;
};
template <>
struct V16K<double> {
  using type = double2;
};

// dot of a matrix row (global or shared) with a shared-memory vector, one warp per row.
template <typename T, bool VEC>
__device__ __forceinline__ T row_dot(const T* __restrict__ row, const T* __restrict__ x, int n,
                                     int lane) {
  T acc = T(0);
  if (VEC) {
    using VT = typename V16K<T>::type;
    constexpr int V = 16 / sizeof(T);
    const VT* r4 = reinterpret_cast<const VT*>(row);
    const VT* x4 = reinterpret_cast<const VT*>(x);
    const int nv = n / V;
    for (int c = lane; c < nv; c += 32) {
      const VT a = r4[c];
      const VT b = x4[c];
      const T* pa = reinterpret_cast<const T*>(&a);
      const T* pb = reinterpret_cast<const T*>(&b);
#pragma unroll
      for (int e = 0; e < V; ++e) acc = fma_(pa[e], pb[e], acc);
    }
  } else {
    for (int c = lane; c < n; c += 32) acc = fma_(row[c], x[c], acc);
  }
  return warp_sum(acc);
}

// y[0:m] = scale * (A[m,n] @ x[0:n]); A row-major with leading dimension lda.
// Each warp takes 4 rows at a time so that every lane keeps >= 8 independent 128-bit loads in
// flight (A streams from HBM/L2 or shared memory); x is read once per 4 rows.
// Ends with __syncthreads().
template <typename T>
__device__ __forceinline__ void cta_matvec(const T* __restrict__ A, int lda, int m, int n,
                                           const T* __restrict__ x, T* __restrict__ y, T scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int V = 16 / sizeof(T);
  constexpr int RB = 4;
  const bool vec = (n % V == 0) && (lda % V == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  if (vec) {
    using VT = typename V16K<T>::type;
    const VT* x4 = reinterpret_cast<const VT*>(x);
    const int nv = n / V;
    for (int i0 = warp * RB; i0 < m; i0 += nw * RB) {
      T acc[RB];
      const VT* rows[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        acc[r] = T(0);
        const int i = i0 + r < m ? i0 + r : m - 1;  // clamp: duplicates are discarded below
        rows[r] = reinterpret_cast<const VT*>(A + (size_t)i * lda);
      }
      for (int c = lane; c < nv; c += 32) {
        VT a[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) a[r] = rows[r][c];
        const VT b = x4[c];
        const T* pb = reinterpret_cast<const T*>(&b);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const T* pa = reinterpret_cast<const T*>(&a[r]);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[r] = fma_(pa[e], pb[e], acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const T s = warp_sum(acc[r]);
        if (lane == 0 && i0 + r < m) y[i0 + r] = scale * s;
      }
    }
  } else {
    for (int i = warp; i < m; i += nw) {
      const T s = row_dot<T, false>(A + (size_t)i * lda, x, n, lane);
      if (lane == 0) y[i] = scale * s;
    }
  }
  __syncthreads();
}

// y[0:n] = scale * (A[m,n]^T @ x[0:m]).  Each warp owns 32 consecutive columns and walks the
// rows; consecutive lanes read consecutive addresses.  Ends with __syncthreads().
template <typename T>
__device__ __forceinline__ void cta_matvec_t(const T* __restrict__ A, int lda, int m, int n,
                                             const T* __restrict__ x, T* __restrict__ y, T scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j0 = warp * 32; j0 < n; j0 += nw * 32) {
    const int j = j0 + lane;
    T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
    if (j < n) {
      int i = 0;
      for (; i + 3 < m; i += 4) {
        acc0 = fma_(A[(size_t)i * lda + j], x[i], acc0);
        acc1 = fma_(A[(size_t)(i + 1) * lda + j], x[i + 1], acc1);
        acc2 = fma_(A[(size_t)(i + 2) * lda + j], x[i + 2], acc2);
        acc3 = fma_(A[(size_t)(i + 3) * lda + j], x[i + 3], acc3);
      }
      for (; i < m; ++i) acc0 = fma_(A[(size_t)i * lda + j], x[i], acc0);
      y[j] = scale * ((acc0 + acc1) + (acc2 + acc3));
    }
  }
  __syncthreads();
}

// Stage a contiguous [rows*cols] matrix into shared memory (16-byte cp.async when possible).
// Ends with __syncthreads().
template <typename T>
__device__ __forceinline__ void cta_stage_matrix(const T* __restrict__ src, T* __restrict__ dst,
                                                 size_t count) {
  constexpr int V = 16 / sizeof(T);
  const bool vec = (count % V == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec) {
    for (size_t c = threadIdx.x; c < count / V; c += blockDim.x) cp_async16(dst + c * V, src + c * V);
    cp_async_commit();
    cp_async_wait<0>();
  } else {
    for (size_t c = threadIdx.x; c < count; c += blockDim.x) dst[c] = src[c];
  }
  __syncthreads();
}

// The lineax convergence test (cg.py:149-160, bicgstab.py:115-126, gmres.py:130-141):
//   not_converged = max_i |r_i / (atol + rtol |b_i|)| > 1  or  max_i |diff_i / (atol + rtol |y_i|)| > 1
// with NaN-propagating max (a NaN norm compares false).  `diff_inf` substitutes diff = +inf.
// All threads return the same answer.  `red` needs 64 elements.
template <typename T>
__device__ __forceinline__ bool cta_not_converged(const T* r, const T* diff, const T* y,
                                                  const T* b, int n, T rtol, T atol, bool has_scale,
                                                  bool diff_inf, T* red) {
  if (!has_scale) return true;
  T v[2] = {T(0), T(0)};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const T bs = atol + rtol * abs_(b[i]);
    const T ys = atol + rtol * abs_(y[i]);
    const T d = diff_inf ? Num<T>::inf() : diff[i];
    v[0] = absmax2(v[0], r[i] / bs);
    v[1] = absmax2(v[1], d / ys);
  }
  block_absmax<T, 2>(v, red);
  return (v[0] > T(1)) || (v[1] > T(1));
}

// When A is streamed on every matvec, limit the number of co-resident systems so that their
// matrices together stay inside the 126 MB L2 (re-reads then hit L2 instead of HBM), but never
// below one CTA per SM.
inline int64_t l2_resident_cap(int64_t cap, size_t mat_bytes) {
  // experiment knob (MB of L2 to budget); 0/unset = no cap.  Measured on B200 (cg256): with
  // the cap the kernel is L2-latency bound and slower than streaming from HBM at full occupancy.
  const char* env = getenv("LXB_L2_CAP_MB");
  const size_t mb = env ? (size_t)atol(env) : 0;
  if (mb == 0) return cap;
  const size_t budget = mb << 20;
  int64_t fit = mat_bytes ? (int64_t)(budget / mat_bytes) : cap;
  if (fit < kNumSMs) fit = kNumSMs;
  return cap < fit ? cap : fit;
}

// cg.py:213-222 (shared by every iterative solver)
__device__ __forceinline__ int krylov_final_result(int64_t steps, int64_t max_steps, int flags,
                                                   bool has_scale) {
  if (!(flags & LXB_MAXSTEPS_GIVEN)) return steps == max_steps ? LXB_SINGULAR : LXB_SUCCESSFUL;
  if (has_scale) return steps == max_steps ? LXB_MAX_STEPS_REACHED : LXB_SUCCESSFUL;
  return LXB_SUCCESSFUL;
}

}  // namespace lxb

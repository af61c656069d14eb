// "Whole grid per system" tier of the Krylov kernels: one persistent COOPERATIVE kernel in which
// every CTA owns a contiguous block of rows of A and the matching slice of every vector.
// Vectors live in global memory (L2-resident), the matvec streams A once from HBM with 128-bit
// non-L1-allocating loads, dot products / norms are reduced through a per-CTA partial array and
// one grid barrier (deterministic order, no atomics), and the whole iteration -- including the
// convergence / breakdown tests -- stays on the device.
#pragma once
#include <cooperative_groups.h>

#include "krylov_cta.cuh"

namespace lxb {

namespace cgx = cooperative_groups;

constexpr int kGridThreads = 256;
constexpr int kGridMaxK = 72;  // widest reduction: QR panel (64), GMRES restart + 2 <= 72

#ifdef LXB_QR_PROF
// debug build only: per-phase nanosecond accumulators of CTA 0 (one copy per translation unit)
static __device__ unsigned long long g_prof[16];
__device__ __forceinline__ unsigned long long prof_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LXB_PROF(slot)                                        \
  do {                                                        \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                \
      const unsigned long long t_ = prof_now();               \
      g_prof[slot] += t_ - g_prof[15];                        \
      g_prof[15] = t_;                                        \
    }                                                         \
  } while (0)
#else
#define LXB_PROF(slot)
#endif

template <typename T>
struct GridTeam {
  cgx::grid_group g;
  int nb, bid, tid, nt;
  T* part;  // global scratch: 2 x kGridMaxK x nb
  int flip;
  T* red;   // shared scratch, >= 96 + kGridMaxK elements

  __device__ GridTeam(T* part_, T* red_)
      : g(cgx::this_grid()), nb(gridDim.x), bid(blockIdx.x), tid(threadIdx.x), nt(blockDim.x),
        part(part_), flip(0), red(red_) {}

  __device__ __forceinline__ void sync() { g.sync(); }

  // contiguous slice [lo, hi) of `n` items owned by this CTA (multiples of 4 where possible)
  __device__ __forceinline__ void slice(int n, int& lo, int& hi) const {
    const int per = (((n + nb - 1) / nb) + 3) & ~3;
    lo = bid * per < n ? bid * per : n;
    hi = lo + per < n ? lo + per : n;
  }

  // All-reduce KS sums and KM NaN-propagating abs-maxima of per-thread values. Every thread of
  // every CTA returns the same bits. Contains exactly one grid barrier.
  template <int KS, int KM>
  __device__ void reduce(T* s, T* m) {
    if (KS > 0) {
      T tmp[KS > 0 ? KS : 1];
      for (int k = 0; k < KS; ++k) tmp[k] = s[k];
      block_sum<T, (KS > 0 ? KS : 1)>(tmp, red);
      for (int k = 0; k < KS; ++k) s[k] = tmp[k];
    }
    if (KM > 0) {
      T tmp[KM > 0 ? KM : 1];
      for (int k = 0; k < KM; ++k) tmp[k] = m[k];
      block_absmax<T, (KM > 0 ? KM : 1)>(tmp, red);
      for (int k = 0; k < KM; ++k) m[k] = tmp[k];
    }
    T* buf = part + (size_t)flip * kGridMaxK * nb;
    flip ^= 1;
    if (tid == 0) {
      for (int k = 0; k < KS; ++k) buf[(size_t)k * nb + bid] = s[k];
      for (int k = 0; k < KM; ++k) buf[(size_t)(KS + k) * nb + bid] = m[k];
      __threadfence();
    }
    g.sync();
    for (int k = 0; k < KS; ++k) {
      T a[1] = {T(0)};
      for (int i = tid; i < nb; i += nt) a[0] += __ldcg(buf + (size_t)k * nb + i);
      block_sum<T, 1>(a, red);
      s[k] = a[0];
    }
    for (int k = 0; k < KM; ++k) {
      T a[1] = {T(0)};
      for (int i = tid; i < nb; i += nt) a[0] = absmax2(a[0], __ldcg(buf + (size_t)(KS + k) * nb + i));
      block_absmax<T, 1>(a, red);
      m[k] = a[0];
    }
  }

  // Runtime-K sum all-reduce: `vals` (shared memory, K <= kGridMaxK entries, already reduced
  // within the CTA and visible to all its threads) -> global sums written back to `vals`.
  __device__ void reduce_dyn(T* vals, int K) {
    T* buf = part + (size_t)flip * kGridMaxK * nb;
    flip ^= 1;
    __syncthreads();
    LXB_PROF(1);
    for (int k = tid; k < K; k += nt) buf[(size_t)k * nb + bid] = vals[k];
    __threadfence();
    LXB_PROF(2);
    g.sync();
    LXB_PROF(3);
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    // a warp sums 8 entries at a time so that their (independent) L2 loads overlap; the order
    // of additions per entry is unchanged: lane-strided partial sums, then the shuffle tree
    constexpr int KW = 8;
    for (int k0 = warp; k0 < K; k0 += nw * KW) {
      T a[KW];
#pragma unroll
      for (int q = 0; q < KW; ++q) a[q] = T(0);
      for (int i = lane; i < nb; i += 32) {
#pragma unroll
        for (int q = 0; q < KW; ++q) {
          const int k = k0 + q * nw;
          if (k < K) a[q] += __ldcg(buf + (size_t)k * nb + i);
        }
      }
#pragma unroll
      for (int q = 0; q < KW; ++q) {
        const int k = k0 + q * nw;
        const T t = warp_sum(a[q]);
        if (k < K && lane == 0) vals[k] = t;
      }
    }
    __syncthreads();
    LXB_PROF(4);
  }
};

static __device__ int g_mv_force_rb = 0;  // per translation unit

// RB rows per warp per sweep, UNR column chunks in flight per row: RB * UNR 128-bit loads per lane.
template <typename T, int RB, int UNR>
__device__ __forceinline__ void grid_matvec_rows(const T* __restrict__ A, int n, int lo, int hi,
                                                 const T* __restrict__ x, T* __restrict__ y, T scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int V = 16 / sizeof(T);
  using VT = typename V16K<T>::type;
  const VT* x4 = reinterpret_cast<const VT*>(x);
  const int nv = n / V;
  for (int i0 = lo + warp * RB; i0 < hi; i0 += nw * RB) {
    T acc[RB];
    const VT* rows[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      acc[r] = T(0);
      const int i = i0 + r < hi ? i0 + r : hi - 1;
      rows[r] = reinterpret_cast<const VT*>(A + (size_t)i * n);
    }
    // explicit UNR-way unroll: RB * UNR independent 128-bit loads are issued before any is used
    int c = lane;
    for (; c + 32 * (UNR - 1) < nv; c += 32 * UNR) {
      VT a[UNR][RB], b[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
#pragma unroll
        for (int r = 0; r < RB; ++r) a[u][r] = ldg_stream(rows[r] + c + 32 * u);
        b[u] = x4[c + 32 * u];
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const T* pb = reinterpret_cast<const T*>(&b[u]);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const T* pa = reinterpret_cast<const T*>(&a[u][r]);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[r] = fma_(pa[e], pb[e], acc[r]);
        }
      }
    }
    for (; c < nv; c += 32) {
      const VT b = x4[c];
      const T* pb = reinterpret_cast<const T*>(&b);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const VT a = ldg_stream(rows[r] + c);
        const T* pa = reinterpret_cast<const T*>(&a);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[r] = fma_(pa[e], pb[e], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const T s = warp_sum(acc[r]);
      if (lane == 0 && i0 + r < hi) y[i0 + r] = scale * s;
    }
  }
}

// y[lo:hi) = scale * A[lo:hi, :] x, A streamed (no L1 allocation), x through L1/L2.
// Ends with __syncthreads().
// When `xs` (shared memory, >= n elements, 16-byte aligned) is given, x is first staged into it
// (one coalesced pass per CTA) and the row dots read it from shared memory: independent of how
// the memory x lives in is cached (peer-mapped symmetric buffers bypass L1).
template <typename T>
__device__ __forceinline__ void grid_matvec(const T* __restrict__ A, int n, int lo, int hi,
                                            const T* __restrict__ x, T* __restrict__ y, T scale,
                                            T* xs = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (xs != nullptr) {
    if ((n % (16 / (int)sizeof(T)) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
      using VTs = typename V16K<T>::type;
      const int nvs = n / (16 / (int)sizeof(T));
      for (int c = threadIdx.x; c < nvs; c += blockDim.x)
        reinterpret_cast<VTs*>(xs)[c] = __ldcg(reinterpret_cast<const VTs*>(x) + c);
    } else {
      for (int c = threadIdx.x; c < n; c += blockDim.x) xs[c] = __ldcg(x + c);
    }
    __syncthreads();
    x = xs;
  }
  constexpr int V = 16 / sizeof(T);
  const bool vec = (n % V == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec) {
    // rows per warp per sweep: keep every warp busy when the CTA owns few rows (multi-GPU slices)
    // Rows per warp per sweep.  Measured on 4 B200s (56 rows per CTA, 16 warps): 4 rows per warp
    // streams ~14 % faster per byte than 2 or 1 (one x chunk feeds 4 rows), so fewer rows per warp
    // only pay when they shorten the critical path (rows the busiest warp has to stream).
    const int rows = hi - lo;
    auto crit = [&](int rbv) { return (rows + nw * rbv - 1) / (nw * rbv) * rbv; };
    int rb = 4, best = crit(4) * 100;
    if (crit(2) * 116 < best) { rb = 2; best = crit(2) * 116; }
    if (crit(1) * 112 < best) { rb = 1; best = crit(1) * 112; }
    if (g_mv_force_rb > 0) rb = g_mv_force_rb;  // experiment knob (LXB_MV_RB), 0 in normal use
    if (rb == 4) grid_matvec_rows<T, 4, 2>(A, n, lo, hi, x, y, scale);
    else if (rb == 2) grid_matvec_rows<T, 2, 4>(A, n, lo, hi, x, y, scale);
    else grid_matvec_rows<T, 1, 8>(A, n, lo, hi, x, y, scale);
  } else {
    for (int i = lo + warp; i < hi; i += nw) {
      const T s = row_dot<T, false>(A + (size_t)i * n, x, n, lane);
      if (lane == 0) y[i] = scale * s;
    }
  }
  __syncthreads();  // y[lo:hi) visible to the whole CTA
}

// Partial A[lo:hi, :]^T u[lo:hi) over this CTA's rows for ALL n columns -> pbuf[bid * n + j].
// Threads own column chunks (coalesced 128-bit loads), loop over the CTA's rows.
template <typename T>
__device__ __forceinline__ void grid_matvec_t_partial(const T* __restrict__ A, int n, int lo, int hi,
                                                      const T* __restrict__ u, T* __restrict__ pout) {
  constexpr int V = 16 / sizeof(T);
  using VT = typename V16K<T>::type;
  const bool vec = (n % V == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(pout) & 15) == 0);
  if (vec) {
    const int nv = n / V;
    for (int c = threadIdx.x; c < nv; c += blockDim.x) {
      T acc[V];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = T(0);
      int i = lo;
      for (; i + 3 < hi; i += 4) {
        VT a[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = ldg_stream(reinterpret_cast<const VT*>(A + (size_t)(i + r) * n) + c);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const T ui = u[i + r];
          const T* pa = reinterpret_cast<const T*>(&a[r]);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = fma_(pa[e], ui, acc[e]);
        }
      }
      for (; i < hi; ++i) {
        const VT a = ldg_stream(reinterpret_cast<const VT*>(A + (size_t)i * n) + c);
        const T ui = u[i];
        const T* pa = reinterpret_cast<const T*>(&a);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = fma_(pa[e], ui, acc[e]);
      }
      VT o;
      T* po = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int e = 0; e < V; ++e) po[e] = acc[e];
      reinterpret_cast<VT*>(pout)[c] = o;
    }
  } else {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      T acc = T(0);
      for (int i = lo; i < hi; ++i) acc = fma_(A[(size_t)i * n + j], u[i], acc);
      pout[j] = acc;
    }
  }
}

}  // namespace lxb

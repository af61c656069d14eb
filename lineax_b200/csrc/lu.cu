// LU with partial pivoting: lineax/_solver/lu.py:43-66 (jsp.linalg.lu_factor / lu_solve).
//
// Tier S (n <= 32): one warp per system, lane r owns row r in registers, the pivot row is
//   broadcast through a 2-deep shared-memory line, A is staged HBM -> smem with 16-byte
//   cp.async into a bank-conflict-free swizzled tile (one 4 KB tile per warp, prefetched
//   for the next system while the current one is being eliminated).
// Tier M (n > 32): one CTA per system, matrix resident in shared memory when it fits
//   (n <= 238 f32 / 168 f64), otherwise factored in place in the caller's `lu` buffer.
//
// Arithmetic (identical in both tiers, and in oracle/getf2.c, so results are bit-identical):
//   right-looking elimination, pivot = first index of max |a_ik| over the not-yet-pivoted
//   rows in LAPACK's physical row order (isamax), r = 1/pivot, l_ik = a_ik * r,
//   a_ij = fma(-l_ik, u_kj, a_ij) for k ascending; solves use y_i = fma(-l_ik, y_k, y_i)
//   and x_k = y_k * (1/u_kk), x_i = fma(-u_ik, x_k, x_i).
#include <stdlib.h>

#include "common.cuh"
#include "lu_tma.cuh"

namespace lxb {

constexpr int kLuWarps = 8;  // warps per CTA in the warp-per-system kernels

template <typename T>
struct V16;
template <>
struct V16<float> {
  using type = float4;
};
template <>
struct V16<double> {
  using type = double2;
};

// swizzled chunk position inside a staged row: conflict-free for 16-byte row reads by
// consecutive lanes (each quarter-warp covers all 32 banks).
template <typename T, int NP>
__device__ __forceinline__ int swz_chunk(int row, int chunk) {
  constexpr int V = 16 / sizeof(T);
  constexpr int CPR = NP / V;                  // 16-byte chunks per row
  constexpr int RB = NP * sizeof(T);           // row bytes
  constexpr int R128 = RB >= 128 ? 1 : 128 / RB;
  constexpr int MASK = (CPR < 8 ? CPR : 8) - 1;
  return chunk ^ ((row / R128) & MASK);
}

template <typename T, int NP>
__device__ __forceinline__ void stage_issue(T* stage, const T* src, int lane) {
  constexpr int V = 16 / sizeof(T);
  constexpr int CPR = NP / V;
#pragma unroll
  for (int c = lane; c < NP * CPR; c += 32) {
    const int row = c / CPR, cc = c % CPR;
    cp_async16(stage + row * NP + swz_chunk<T, NP>(row, cc) * V, src + (size_t)c * V);
  }
  cp_async_commit();
}

template <typename T, int NP>
__device__ __forceinline__ void stage_read_row(const T* stage, int row, T (&a)[NP]) {
  using VT = typename V16<T>::type;
  constexpr int V = 16 / sizeof(T);
  constexpr int CPR = NP / V;
#pragma unroll
  for (int c = 0; c < CPR; ++c) {
    VT v = *reinterpret_cast<const VT*>(stage + row * NP + swz_chunk<T, NP>(row, c) * V);
    const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int e = 0; e < V; ++e) a[c * V + e] = pv[e];
  }
}

// Pivot choice: max |a| over candidate lanes, ties -> smallest physical position (LAPACK ISAMAX).
// Returns a PREDICATE (am I the pivot lane) instead of a lane index: ncu showed the XU pipe
// (FLO/BREV/POPC from __ffs/__popc, MUFU) at 79 % -- the real limiter of the first version -- so the
// hot loop avoids index extraction altogether and broadcasts through shared memory.
__device__ __forceinline__ bool pick_pivot(float v, bool cand, int pos) {
  // every NaN maps to one key so that, like ISAMAX, the first NaN wins
  const unsigned key = cand ? (v != v ? 0x7fc00000u : (__float_as_uint(v) & 0x7fffffffu)) + 1u : 0u;
  const unsigned m = __reduce_max_sync(kFull, key);
  bool is = key == m;
  const unsigned tied = __ballot_sync(kFull, is);
  if (tied & (tied - 1u)) {  // more than one candidate holds the maximum
    const int pm = __reduce_min_sync(kFull, is ? pos : 1 << 20);
    is = is && pos == pm;
  }
  return is;
}
__device__ __forceinline__ bool pick_pivot(double v, bool cand, int pos) {
  const unsigned long long key =
      cand ? (v != v ? 0x7ff8000000000000ull
                     : ((unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull)) + 1ull
           : 0ull;
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
  bool is = hi == mh && lo == ml;
  const unsigned tied = __ballot_sync(kFull, is);
  if (tied & (tied - 1u)) {
    const int pm = __reduce_min_sync(kFull, is ? pos : 1 << 20);
    is = is && pos == pm;
  }
  return is;
}

__device__ __forceinline__ float int_as_T(int v, float) { return __int_as_float(v); }
__device__ __forceinline__ double int_as_T(int v, double) { return __longlong_as_double((long long)v); }
__device__ __forceinline__ int T_as_int(float v) { return __float_as_int(v); }
__device__ __forceinline__ int T_as_int(double v) { return (int)__double_as_longlong(v); }

constexpr int kLuLineExtra = 4;  // per broadcast line: 1/pivot, pivot-row rhs entry, pivot position, pad

// Elimination on a register-resident system. On exit lane r holds (in LAPACK's final
// layout) row `pos` of the packed LU factors, `bb` the forward-substituted right-hand side
// for that row, `rdiag` = 1/u_pos,pos and `mypiv` = piv[lane].
template <typename T, int NP, bool SOLVE, bool FULL>
__device__ __forceinline__ void lu_eliminate(T (&a)[NP], T& bb, int& pos, T& rdiag, int& mypiv,
                                             T* urow, int lane, int n) {
  using VT = typename V16<T>::type;
  constexpr int V = 16 / sizeof(T);
  constexpr int CPR = NP / V;
  const bool real = NP == 32 || lane < NP;
  constexpr int LINE = NP + kLuLineExtra;
  pos = lane;
  rdiag = T(0);
  mypiv = lane;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    if (FULL || k < n) {
      // every lane inverts its own candidate while the pivot search is in flight; the pivot
      // lane's value is the 1/pivot everybody needs (same IEEE division, shorter critical path)
      const T rown = T(1) / a[k];
      const bool is_pl = pick_pivot(a[k], real && pos >= k, pos);
      T* ub = urow + (k & 1) * LINE;
      if (is_pl) {
#pragma unroll
        for (int c = k / V; c < CPR; ++c) {
          VT v;
          T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
          for (int e = 0; e < V; ++e) pv[e] = a[c * V + e];
          *reinterpret_cast<VT*>(ub + c * V) = v;
        }
        T ex[kLuLineExtra] = {rown, bb, int_as_T(pos, T(0)), T(0)};
#pragma unroll
        for (int c = 0; c < kLuLineExtra / V; ++c) {
          VT v;
          T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
          for (int e = 0; e < V; ++e) pv[e] = ex[c * V + e];
          *reinterpret_cast<VT*>(ub + NP + c * V) = v;
        }
      }
      __syncwarp();
      T u[NP];
#pragma unroll
      for (int c = k / V; c < CPR; ++c) {
        VT v = *reinterpret_cast<const VT*>(ub + c * V);
        const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int e = 0; e < V; ++e) u[c * V + e] = pv[e];
      }
      T ex[kLuLineExtra];
#pragma unroll
      for (int c = 0; c < kLuLineExtra / V; ++c) {
        VT v = *reinterpret_cast<const VT*>(ub + NP + c * V);
        const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int e = 0; e < V; ++e) ex[c * V + e] = pv[e];
      }
      const T r = ex[0], bpk = ex[1];
      const int ppos = T_as_int(ex[2]);  // physical position of the pivot row = piv[k]
      if (lane == k) mypiv = ppos;
      if (pos == k) pos = ppos;
      if (is_pl) {
        pos = k;
        rdiag = r;
      }
      if (real && pos > k) {
        const T l = a[k] * r;
        a[k] = l;
#pragma unroll
        for (int j = k + 1; j < NP; ++j) a[j] = fma_(-l, u[j], a[j]);
        if (SOLVE) bb = fma_(-l, bpk, bb);
      }
    }
  }
}

// Back substitution U x = y on the register-resident factors (rows addressed by `pos`).
template <typename T, int NP, bool FULL>
__device__ __forceinline__ T lu_backsolve(const T (&a)[NP], T bb, int pos, T rdiag, int lane,
                                          int n, T* slots /* 2 elements of shared memory */) {
  T mine = T(0);
#pragma unroll
  for (int k = NP - 1; k >= 0; --k) {
    if (FULL || k < n) {
      if (pos == k) slots[k & 1] = bb * rdiag;  // owner of final row k publishes x_k
      __syncwarp();
      const T xk = slots[k & 1];
      if (pos < k) bb = fma_(-a[k], xk, bb);
      if (lane == k) mine = xk;
    }
  }
  return mine;
}

// MODE: 0 = factor only, 1 = factor + solve (lu/piv optional)
template <typename T, int NP, bool SOLVE, bool FULL, int MINB = ((sizeof(T) == 4 && NP == 32) ? 3 : 1)>
__global__ void __launch_bounds__(kLuWarps * 32, MINB)
    lu_warp_kernel(const T* __restrict__ A, int64_t sA, const T* __restrict__ B, int64_t sB,
                   T* __restrict__ X, T* __restrict__ LU, int32_t* __restrict__ PIV, int64_t batch,
                   int n, int fast) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* stage = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (NP * NP + 2 * (NP + kLuLineExtra));
  T* urow = stage + NP * NP;
  const int64_t nwarps = (int64_t)gridDim.x * kLuWarps;
  int64_t sys = (int64_t)blockIdx.x * kLuWarps + warp;
  if (fast && sys < batch) stage_issue<T, NP>(stage, A + sys * sA, lane);
  for (; sys < batch; sys += nwarps) {
    T a[NP];
    if (fast) {
      cp_async_wait<0>();
      __syncwarp();
      if (lane < NP) stage_read_row<T, NP>(stage, lane, a);
      __syncwarp();
      const int64_t nxt = sys + nwarps;
      if (nxt < batch) stage_issue<T, NP>(stage, A + nxt * sA, lane);
    } else {
      const T* src = A + sys * sA;
#pragma unroll
      for (int j = 0; j < NP; ++j)
        a[j] = (lane < n && j < n) ? src[(size_t)lane * n + j] : (lane == j ? T(1) : T(0));
    }
    T bb = T(0);
    if (SOLVE && lane < n) bb = B[sys * sB + lane];
    int pos, mypiv;
    T rdiag;
    lu_eliminate<T, NP, SOLVE, FULL>(a, bb, pos, rdiag, mypiv, urow, lane, n);
    if (LU != nullptr && lane < NP && pos < n) {
      T* dst = LU + sys * (int64_t)n * n + (size_t)pos * n;
      if (fast) {
        using VT = typename V16<T>::type;
        constexpr int V = 16 / sizeof(T);
#pragma unroll
        for (int c = 0; c < NP / V; ++c) {
          VT v;
          T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
          for (int e = 0; e < V; ++e) pv[e] = a[c * V + e];
          reinterpret_cast<VT*>(dst)[c] = v;
        }
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (j < n) dst[j] = a[j];
      }
    }
    if (PIV != nullptr && lane < n) PIV[sys * n + lane] = mypiv;
    if (SOLVE) {
      __syncwarp();
      const T xm = lu_backsolve<T, NP, FULL>(a, bb, pos, rdiag, lane, n, urow);
      if (lane < n) X[sys * n + lane] = xm;
    }
  }
}

// x = lu_solve((lu, piv), b, trans): lane r holds row r of LU (or of LU^T when TRANS).
template <typename T, int NP, bool TRANS>
__global__ void __launch_bounds__(kLuWarps * 32)
    lu_solve_warp_kernel(const T* __restrict__ LU, int64_t sLU, const int32_t* __restrict__ PIV,
                         int64_t sP, const T* __restrict__ B, int64_t sB, T* __restrict__ X,
                         int64_t batch, int n) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * kLuWarps;
  for (int64_t sys = (int64_t)blockIdx.x * kLuWarps + warp; sys < batch; sys += nwarps) {
    const T* src = LU + sys * sLU;
    T a[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const bool in = lane < n && j < n;
      a[j] = in ? (TRANS ? src[(size_t)j * n + lane] : src[(size_t)lane * n + j])
                : (lane == j ? T(1) : T(0));
    }
    const T rdiag = lane < n ? T(1) / src[(size_t)lane * n + lane] : T(1);
    const int mypiv = lane < n ? PIV[sys * sP + lane] : lane;
    T y = lane < n ? B[sys * sB + lane] : T(0);
    if (!TRANS) {
      // y = P b : apply the row swaps in order (getrs, trans = 'N')
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        if (k < n) {
          const int pk = __shfl_sync(kFull, mypiv, k);
          const T yk = __shfl_sync(kFull, y, k), yp = __shfl_sync(kFull, y, pk);
          if (lane == k) y = yp;
          else if (lane == pk) y = yk;
        }
      }
      // L y' = y (unit lower)
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        if (k < n) {
          const T yk = __shfl_sync(kFull, y, k);
          if (lane > k) y = fma_(-a[k], yk, y);
        }
      }
      // U x = y'
#pragma unroll
      for (int k = NP - 1; k >= 0; --k) {
        if (k < n) {
          const T xk = __shfl_sync(kFull, y * rdiag, k);
          if (lane < k) y = fma_(-a[k], xk, y);
          else if (lane == k) y = xk;
        }
      }
    } else {
      // a = (LU)^T: lower part (incl. diagonal) = U^T, strict upper part = L^T (unit)
      // U^T z = b
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        if (k < n) {
          const T zk = __shfl_sync(kFull, y * rdiag, k);
          if (lane > k) y = fma_(-a[k], zk, y);
          else if (lane == k) y = zk;
        }
      }
      // L^T w = z
#pragma unroll
      for (int k = NP - 1; k >= 0; --k) {
        if (k < n) {
          const T wk = __shfl_sync(kFull, y, k);
          if (lane < k) y = fma_(-a[k], wk, y);
        }
      }
      // x = P^T w : undo the swaps in reverse order
#pragma unroll
      for (int k = NP - 1; k >= 0; --k) {
        if (k < n) {
          const int pk = __shfl_sync(kFull, mypiv, k);
          const T yk = __shfl_sync(kFull, y, k), yp = __shfl_sync(kFull, y, pk);
          if (lane == k) y = yp;
          else if (lane == pk) y = yk;
        }
      }
    }
    if (lane < n) X[sys * n + lane] = y;
  }
}

// ---------------------------------------------------------------- Tier M ----
constexpr int kLuBlockThreads = 256;

// One CTA per system. M points at the working matrix (shared or global), leading dim ld.
template <typename T>
__device__ void lu_block_factor(T* M, int ld, int n, T* y /* rhs or null */, int* piv_s, T* red_v,
                                int* red_i) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int k = 0; k < n; ++k) {
    // --- pivot search: first index of max |M[i][k]|, i >= k
    T bv = T(-1);
    int bi = n;
    for (int i = k + tid; i < n; i += nt) {
      const T v = abs_(M[(size_t)i * ld + k]);
      if (v > bv || (v != v && bv == bv)) {  // strict > keeps the first index; NaN wins once
        bv = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(kFull, bv, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      const bool take = (ov > bv) || (ov == bv && oi < bi) || (ov != ov && (bv == bv || oi < bi));
      if (take) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      red_v[warp] = bv;
      red_i[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
      bv = lane < nw ? red_v[lane] : T(-1);
      bi = lane < nw ? red_i[lane] : n;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(kFull, bv, o);
        const int oi = __shfl_xor_sync(kFull, bi, o);
        const bool take = (ov > bv) || (ov == bv && oi < bi) || (ov != ov && (bv == bv || oi < bi));
        if (take) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) piv_s[k] = bi < n ? bi : k;
    }
    __syncthreads();
    const int p = piv_s[k];
    // --- swap rows k and p (whole rows, LAPACK laswp semantics)
    if (p != k) {
      for (int j = tid; j < n; j += nt) {
        const T t = M[(size_t)k * ld + j];
        M[(size_t)k * ld + j] = M[(size_t)p * ld + j];
        M[(size_t)p * ld + j] = t;
      }
      if (y != nullptr && tid == 0) {
        const T t = y[k];
        y[k] = y[p];
        y[p] = t;
      }
    }
    __syncthreads();
    const T r = T(1) / M[(size_t)k * ld + k];
    // --- scale column k, rank-1 update of the trailing block (and of the rhs)
    const int rem = n - k - 1;
    for (int idx = tid; idx < rem * (rem + 1); idx += nt) {
      // rem rows x (rem + 1) columns: column 0 is the multiplier slot, handled after
      const int i = k + 1 + idx / (rem + 1), jj = idx % (rem + 1);
      if (jj == 0) continue;
      const int j = k + jj;
      const T l = M[(size_t)i * ld + k] * r;
      M[(size_t)i * ld + j] = fma_(-l, M[(size_t)k * ld + j], M[(size_t)i * ld + j]);
    }
    if (y != nullptr) {
      for (int i = k + 1 + tid; i < n; i += nt) y[i] = fma_(-(M[(size_t)i * ld + k] * r), y[k], y[i]);
    }
    __syncthreads();
    for (int i = k + 1 + tid; i < n; i += nt) M[(size_t)i * ld + k] *= r;
    __syncthreads();
  }
}

// U x = y in place in y; M holds packed LU.
template <typename T>
__device__ void block_upper_solve(const T* M, int ld, int n, T* y) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = n - 1; k >= 0; --k) {
    if (tid == 0) y[k] = y[k] * (T(1) / M[(size_t)k * ld + k]);
    __syncthreads();
    const T xk = y[k];
    for (int i = tid; i < k; i += nt) y[i] = fma_(-M[(size_t)i * ld + k], xk, y[i]);
    __syncthreads();
  }
}

template <typename T, bool SOLVE>
__global__ void __launch_bounds__(kLuBlockThreads)
    lu_block_kernel(const T* __restrict__ A, int64_t sA, const T* __restrict__ B, int64_t sB,
                    T* __restrict__ X, T* __restrict__ LU, int32_t* __restrict__ PIV, int64_t batch,
                    int n, int in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  // smem: [matrix n*(n+1) if in_smem] [y n] [red_v 32] [piv n ints] [red_i 32 ints]
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int ld = in_smem ? n + 1 : n;
  T* Ms = sm;
  T* y = sm + (in_smem ? (size_t)n * (n + 1) : 0);
  T* red_v = y + n;
  int* piv_s = reinterpret_cast<int*>(red_v + 32);
  int* red_i = piv_s + n;
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    const T* src = A + sys * sA;
    T* M = in_smem ? Ms : LU + sys * (int64_t)n * n;
    for (int idx = tid; idx < n * n; idx += nt) M[(size_t)(idx / n) * ld + idx % n] = src[idx];
    if (SOLVE)
      for (int i = tid; i < n; i += nt) y[i] = B[sys * sB + i];
    __syncthreads();
    lu_block_factor<T>(M, ld, n, SOLVE ? y : nullptr, piv_s, red_v, red_i);
    if (SOLVE) {
      block_upper_solve<T>(M, ld, n, y);
      for (int i = tid; i < n; i += nt) X[sys * n + i] = y[i];
    }
    if (in_smem && LU != nullptr) {
      T* dst = LU + sys * (int64_t)n * n;
      for (int idx = tid; idx < n * n; idx += nt) dst[idx] = M[(size_t)(idx / n) * ld + idx % n];
    }
    if (PIV != nullptr)
      for (int i = tid; i < n; i += nt) PIV[sys * n + i] = piv_s[i];
    __syncthreads();
  }
}

// lu_solve for n > 32: one CTA per system, factors read from global (L2), vector in smem.
template <typename T>
__global__ void __launch_bounds__(kLuBlockThreads)
    lu_solve_block_kernel(const T* __restrict__ LU, int64_t sLU, const int32_t* __restrict__ PIV,
                          int64_t sP, const T* __restrict__ B, int64_t sB, T* __restrict__ X,
                          int64_t batch, int n, int trans) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* y = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int64_t sys = blockIdx.x; sys < batch; sys += gridDim.x) {
    const T* M = LU + sys * sLU;
    const int32_t* piv = PIV + sys * sP;
    for (int i = tid; i < n; i += nt) y[i] = B[sys * sB + i];
    __syncthreads();
    if (!trans) {
      if (tid == 0)
        for (int k = 0; k < n; ++k) {
          const int p = piv[k];
          const T t = y[k];
          y[k] = y[p];
          y[p] = t;
        }
      __syncthreads();
      for (int k = 0; k < n; ++k) {
        const T yk = y[k];
        for (int i = k + 1 + tid; i < n; i += nt) y[i] = fma_(-M[(size_t)i * n + k], yk, y[i]);
        __syncthreads();
      }
      for (int k = n - 1; k >= 0; --k) {
        if (tid == 0) y[k] = y[k] * (T(1) / M[(size_t)k * n + k]);
        __syncthreads();
        const T xk = y[k];
        for (int i = tid; i < k; i += nt) y[i] = fma_(-M[(size_t)i * n + k], xk, y[i]);
        __syncthreads();
      }
    } else {
      for (int k = 0; k < n; ++k) {  // U^T z = b
        if (tid == 0) y[k] = y[k] * (T(1) / M[(size_t)k * n + k]);
        __syncthreads();
        const T zk = y[k];
        for (int i = k + 1 + tid; i < n; i += nt) y[i] = fma_(-M[(size_t)k * n + i], zk, y[i]);
        __syncthreads();
      }
      for (int k = n - 1; k >= 0; --k) {  // L^T w = z
        const T wk = y[k];
        for (int i = tid; i < k; i += nt) y[i] = fma_(-M[(size_t)k * n + i], wk, y[i]);
        __syncthreads();
      }
      if (tid == 0)
        for (int k = n - 1; k >= 0; --k) {
          const int p = piv[k];
          const T t = y[k];
          y[k] = y[p];
          y[p] = t;
        }
      __syncthreads();
    }
    for (int i = tid; i < n; i += nt) X[sys * n + i] = y[i];
    __syncthreads();
  }
}

// ------------------------------------------------------------ launchers ----
constexpr size_t kMaxSmem = 227 * 1024;

template <typename T, int NP, bool SOLVE, bool FULL>
int launch_lu_warp_impl(const T* A, int64_t sA, const T* b, int64_t sb, T* x, T* lu, int32_t* piv,
                        int64_t batch, int n, cudaStream_t st) {
  const size_t smem = (size_t)kLuWarps * (NP * NP + 2 * (NP + kLuLineExtra)) * sizeof(T);
  auto kern = lu_warp_kernel<T, NP, SOLVE, FULL>;
  if constexpr (sizeof(T) == 4 && NP == 32 && SOLVE && FULL) {
    const char* e = getenv("LXB_LU_MINB");  // tuning knob: resident CTAs per SM the compiler targets
    if (e && atoi(e) == 4) kern = lu_warp_kernel<T, NP, SOLVE, FULL, 4>;
    if (e && atoi(e) == 2) kern = lu_warp_kernel<T, NP, SOLVE, FULL, 2>;
  }
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int fast = (n == NP) && aligned16(A) && ((sA * sizeof(T)) % 16 == 0) &&
                   (lu == nullptr || aligned16(lu));
  int occ = 1;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kLuWarps * 32, smem));
  if (occ < 1) occ = 1;
  int64_t blocks = (batch + kLuWarps - 1) / kLuWarps;
  const int64_t cap = (int64_t)kNumSMs * occ;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, kLuWarps * 32, smem, st>>>(A, sA, b, sb, x, lu, piv, batch, n, fast);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T, int NP, bool SOLVE>
int launch_lu_warp(const T* A, int64_t sA, const T* b, int64_t sb, T* x, T* lu, int32_t* piv,
                   int64_t batch, int n, cudaStream_t st) {
  if (n == NP) return launch_lu_warp_impl<T, NP, SOLVE, true>(A, sA, b, sb, x, lu, piv, batch, n, st);
  return launch_lu_warp_impl<T, NP, SOLVE, false>(A, sA, b, sb, x, lu, piv, batch, n, st);
}

template <typename T, bool SOLVE>
int lu_dispatch(const T* A, int64_t sA, const T* b, int64_t sb, T* x, T* lu, int32_t* piv,
                int64_t batch, int n, cudaStream_t st) {
  if (batch < 0 || n < 0 || A == nullptr) return LXB_E_BADARG;
  if (SOLVE && (b == nullptr || x == nullptr)) return LXB_E_BADARG;
  if (!SOLVE && (lu == nullptr || piv == nullptr)) return LXB_E_BADARG;
  if (batch == 0 || n == 0) return 0;
  if constexpr (sizeof(T) == 4) {
    // BASELINE configs[1]: TMA-staged FFMA2 kernel (lu_tma.cu); LXB_LU_LEGACY=1 keeps the cp.async one
    static const bool legacy = getenv("LXB_LU_LEGACY") != nullptr;
    if (!legacy && lu32_tma_eligible(A, sA, lu, batch, n))
      return lu32_tma_launch(A, sA, b, sb, x, lu, piv, batch, SOLVE, st);
  }
  if (n <= 8) return launch_lu_warp<T, 8, SOLVE>(A, sA, b, sb, x, lu, piv, batch, n, st);
  if (n <= 16) return launch_lu_warp<T, 16, SOLVE>(A, sA, b, sb, x, lu, piv, batch, n, st);
  if (n <= 32) return launch_lu_warp<T, 32, SOLVE>(A, sA, b, sb, x, lu, piv, batch, n, st);
  // Tier M
  const size_t vec_bytes = ((size_t)n + 32) * sizeof(T) + ((size_t)n + 32) * sizeof(int);
  const size_t mat_bytes = (size_t)n * (n + 1) * sizeof(T);
  const int in_smem = mat_bytes + vec_bytes <= kMaxSmem;
  if (!in_smem && lu == nullptr) return LXB_E_WORKSPACE;  // need the lu buffer as workspace
  if (vec_bytes > kMaxSmem) return LXB_E_UNSUPPORTED;
  const size_t smem = (in_smem ? mat_bytes : 0) + vec_bytes;
  auto kern = lu_block_kernel<T, SOLVE>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kLuBlockThreads, smem));
  if (occ < 1) occ = 1;
  int64_t blocks = batch < (int64_t)kNumSMs * occ ? batch : (int64_t)kNumSMs * occ;
  kern<<<(unsigned)blocks, kLuBlockThreads, smem, st>>>(A, sA, b, sb, x, lu, piv, batch, n, in_smem);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T, int NP>
int launch_lu_solve_warp(const T* lu, int64_t sLU, const int32_t* piv, int64_t sP, const T* b,
                         int64_t sb, T* x, int64_t batch, int n, int trans, cudaStream_t st) {
  int64_t blocks = (batch + kLuWarps - 1) / kLuWarps;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (trans)
    lu_solve_warp_kernel<T, NP, true>
        <<<(unsigned)blocks, kLuWarps * 32, 0, st>>>(lu, sLU, piv, sP, b, sb, x, batch, n);
  else
    lu_solve_warp_kernel<T, NP, false>
        <<<(unsigned)blocks, kLuWarps * 32, 0, st>>>(lu, sLU, piv, sP, b, sb, x, batch, n);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int lu_solve_dispatch(const T* lu, int64_t sLU, const int32_t* piv, int64_t sP, const T* b,
                      int64_t sb, T* x, int64_t batch, int n, int flags, cudaStream_t st) {
  if (batch < 0 || n < 0 || lu == nullptr || piv == nullptr || b == nullptr || x == nullptr)
    return LXB_E_BADARG;
  if (batch == 0 || n == 0) return 0;
  const int trans = (flags & LXB_TRANS) ? 1 : 0;
  if (n <= 8) return launch_lu_solve_warp<T, 8>(lu, sLU, piv, sP, b, sb, x, batch, n, trans, st);
  if (n <= 16) return launch_lu_solve_warp<T, 16>(lu, sLU, piv, sP, b, sb, x, batch, n, trans, st);
  if (n <= 32) return launch_lu_solve_warp<T, 32>(lu, sLU, piv, sP, b, sb, x, batch, n, trans, st);
  const size_t smem = (size_t)n * sizeof(T);
  if (smem > kMaxSmem) return LXB_E_UNSUPPORTED;
  auto kern = lu_solve_block_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = batch < (int64_t)kNumSMs * 4 ? batch : (int64_t)kNumSMs * 4;
  kern<<<(unsigned)blocks, kLuBlockThreads, smem, st>>>(lu, sLU, piv, sP, b, sb, x, batch, n, trans);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_LU(sfx, T)                                                                       \
  extern "C" int lxb_lu_factor_##sfx(const T* A, int64_t stride_A, T* lu, int32_t* piv,          \
                                     int64_t batch, int32_t n, lxb_stream_t stream) {            \
    return lxb::lu_dispatch<T, false>(A, stride_A, nullptr, 0, nullptr, lu, piv, batch, n,       \
                                      (cudaStream_t)stream);                                     \
  }                                                                                              \
  extern "C" int lxb_lu_solve_##sfx(const T* lu, int64_t stride_lu, const int32_t* piv,          \
                                    int64_t stride_piv, const T* b, int64_t stride_b, T* x,      \
                                    int64_t batch, int32_t n, int32_t flags,                     \
                                    lxb_stream_t stream) {                                       \
    return lxb::lu_solve_dispatch<T>(lu, stride_lu, piv, stride_piv, b, stride_b, x, batch, n,   \
                                     flags, (cudaStream_t)stream);                               \
  }                                                                                              \
  extern "C" int lxb_lu_factor_solve_##sfx(const T* A, int64_t stride_A, const T* b,             \
                                           int64_t stride_b, T* x, T* lu, int32_t* piv,          \
                                           int64_t batch, int32_t n, lxb_stream_t stream) {      \
    return lxb::lu_dispatch<T, true>(A, stride_A, b, stride_b, x, lu, piv, batch, n,             \
                                     (cudaStream_t)stream);                                      \
  }
LXB_DEF_LU(f32, float)
LXB_DEF_LU(f64, double)

// Device helpers shared by the single-GPU grid LSMR kernel (krylov_grid.cu) and the row-sharded
// multi-GPU one (lsmr_dist.cu): deterministic grid dot / norm, the SciPy-style Givens rotation of
// lineax/_solver/lsmr.py:361-409 and the one-read-of-A Golub-Kahan pass.
#pragma once
#include "krylov_grid.cuh"

namespace lxb {

template <typename T>
__device__ __forceinline__ T grid_dot(GridTeam<T>& team, const T* a, const T* b, int lo, int hi) {
  T s[1] = {T(0)};
  for (int i = lo + team.tid; i < hi; i += team.nt) s[0] = fma_(a[i], b[i], s[0]);
  team.template reduce<1, 0>(s, nullptr);
  return s[0];
}

// two_norm over a distributed vector of total length n (size-1 shortcut, _norm.py:74-80)
template <typename T>
__device__ __forceinline__ T grid_norm2(GridTeam<T>& team, const T* a, int lo, int hi, int n) {
  if (n == 1) {
    team.sync();
    return abs_(a[0]);
  }
  return sqrt_(grid_dot<T>(team, a, a, lo, hi));
}

// ---------------------------------------------------------------------- LSMR ----
template <typename T>
__device__ __forceinline__ T sign_g(T a) {
  return a > T(0) ? T(1) : (a < T(0) ? T(-1) : a);
}
template <typename T>
__device__ __forceinline__ void givens_g(T a, T b, T& c, T& s, T& r) {
  if (a == T(0) || b == T(0)) {
    if (b == T(0)) {
      c = sign_g(a); s = T(0); r = abs_(a);
    } else {
      c = T(0); s = sign_g(b); r = abs_(b);
    }
  } else if (abs_(b) > abs_(a)) {
    const T tau = a / b;
    s = sign_g(b) / sqrt_(T(1) + tau * tau);
    c = s * tau;
    r = b / (s == T(0) ? T(1) : s);
  } else {
    const T tau = b / a;
    c = sign_g(a) / sqrt_(T(1) + tau * tau);
    s = c * tau;
    r = a / (c == T(0) ? T(1) : c);
  }
}

// out[clo:chi) = out * scale_old + (sum over CTAs of the A^T u partials), fixed summation order:
// one warp per column, lanes stride over the CTAs' partial rows.
template <typename T>
__device__ __forceinline__ void grid_reduce_cols(const GridTeam<T>& team, const T* pbuf, int n,
                                                 int clo, int chi, T* out_scaled_add, T scale_old) {
  const int lane = team.tid & 31, warp = team.tid >> 5, nw = team.nt >> 5;
  for (int j = clo + warp; j < chi; j += nw) {
    T acc = T(0);
    for (int b = lane; b < team.nb; b += 32) acc += __ldcg(pbuf + (size_t)b * n + j);
    acc = warp_sum(acc);
    if (lane == 0) out_scaled_add[j] = out_scaled_add[j] * scale_old + acc;
  }
}

// Single pass over this CTA's rows of A that produces BOTH Golub-Kahan products:
//   u'_i   = s1 * (A_i . vin) + s2 * uold_i            (written to unew, rows rlo..rhi)
//   pout_j = sum_i A_ij u'_i                            (this CTA's partial of A^T u')
// so A is read from HBM exactly once per LSMR iteration (the reference reads it twice,
// lsmr.py:214-237; normalising by beta = ||u'|| commutes with the second product).
// Thread t owns the 16-byte column chunks t, t+256, ... (CH of them) in registers for 4 rows.
// Returns this CTA's sum of u'_i^2 (identical in every thread).
template <typename T, int CH>
__device__ __forceinline__ T lsmr_fused_pass(const T* __restrict__ A, int n, int rlo, int rhi,
                                             const T* vin, const T* uold, T* unew /* may alias uold */,
                                             T s1, T s2, T* __restrict__ pout,
                                             T* red) {
  constexpr int V = 16 / sizeof(T);
  constexpr int R = 4;
  using VT = typename V16K<T>::type;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nv = n / V;
  VT v[CH];
  T acc[CH][V];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = tid + k * nt;
    if (c < nv) {
      v[k] = reinterpret_cast<const VT*>(vin)[c];
    } else {
      T* pz = reinterpret_cast<T*>(&v[k]);
#pragma unroll
      for (int e = 0; e < V; ++e) pz[e] = T(0);
    }
#pragma unroll
    for (int e = 0; e < V; ++e) acc[k][e] = T(0);
  }
  T ssq = T(0);
  for (int i0 = rlo; i0 < rhi; i0 += R) {
    VT a[R][CH];
    T part[R], uo[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = i0 + r < rhi ? i0 + r : rhi - 1;
      // unew may alias uold: every thread takes its copy BEFORE thread 0 overwrites the entry below
      // (the block reduction in between is the barrier)
      uo[r] = uold[i];
      const VT* row = reinterpret_cast<const VT*>(A + (size_t)i * n);
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = tid + k * nt;
        if (c < nv) a[r][k] = ldg_stream(row + c);
        else a[r][k] = v[k];  // v[k] is zero there: contributes nothing
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      T d = T(0);
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const T* pa = reinterpret_cast<const T*>(&a[r][k]);
        const T* pv = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
        for (int e = 0; e < V; ++e) d = fma_(pa[e], pv[e], d);
      }
      part[r] = d;
    }
    block_sum<T, R>(part, red);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool valid = i0 + r < rhi;
      T un = T(0);
      if (valid) {
        un = s1 * part[r] + s2 * uo[r];
        if (tid == 0) unew[i0 + r] = un;
        ssq = fma_(un, un, ssq);
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const T* pa = reinterpret_cast<const T*>(&a[r][k]);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[k][e] = fma_(pa[e], un, acc[k][e]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = tid + k * nt;
    if (c < nv) {
      VT o;
      T* po = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int e = 0; e < V; ++e) po[e] = acc[k][e];
      reinterpret_cast<VT*>(pout)[c] = o;
    }
  }
  return ssq;
}

template <typename T>
__device__ __forceinline__ T lsmr_fused_dispatch(int ch, const T* A, int n, int rlo, int rhi,
                                                 const T* vin, const T* uold, T* unew, T s1, T s2,
                                                 T* pout, T* red) {
  switch (ch) {
    case 1: return lsmr_fused_pass<T, 1>(A, n, rlo, rhi, vin, uold, unew, s1, s2, pout, red);
    case 2: return lsmr_fused_pass<T, 2>(A, n, rlo, rhi, vin, uold, unew, s1, s2, pout, red);
    case 3: return lsmr_fused_pass<T, 3>(A, n, rlo, rhi, vin, uold, unew, s1, s2, pout, red);
    default: return lsmr_fused_pass<T, 4>(A, n, rlo, rhi, vin, uold, unew, s1, s2, pout, red);
  }
}

// out[clo:chi) = (sum over CTAs of partials) * inv + out * scale_old   (v update after the fused pass)
template <typename T>
__device__ __forceinline__ void grid_reduce_cols_scaled(const GridTeam<T>& team, const T* pbuf, int n,
                                                        int clo, int chi, T* out, T inv_div,
                                                        T scale_old) {
  const int lane = team.tid & 31, warp = team.tid >> 5, nw = team.nt >> 5;
  for (int j = clo + warp; j < chi; j += nw) {
    T acc = T(0);
    for (int b = lane; b < team.nb; b += 32) acc += __ldcg(pbuf + (size_t)b * n + j);
    acc = warp_sum(acc);
    if (lane == 0) out[j] = out[j] * scale_old + acc / inv_div;
  }
}

}  // namespace lxb

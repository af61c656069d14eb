// Blocked Householder QR (compact WY, LAPACK geqrf/larft/larfb conventions) for ONE large tall
// matrix on all SMs -- BASELINE configs[4] QR (262144 x 4096 fp32) -- plus the matching
// Q^T b application for the least-squares solve (lineax/_solver/qr.py:55-94).
//
//   for each panel of 32 columns:
//     K1 qr_panel_kernel   (cooperative): each CTA keeps its rows of the panel in shared memory;
//                          per column ONE grid all-reduce delivers the Gram row (column norm and
//                          all v^T a_c at once) and the pivot row; then V^T V -> T (larft).
//     K2 qr_wpartial_kernel: Wp[g] = V[rows_g]^T A2[rows_g]   (register-tiled, row groups)
//     K3 qr_wfinish_kernel : W2 = T^T (sum_g Wp[g])
//     K4 qr_update_kernel  : A2 -= V W2                        (register-tiled 64x128 tiles)
// Everything is fp32/fp64 SIMT FMA: the 1e-5 parity budget rules out plain TF32 tensor-core
// tiles (SURVEY.md section 7 "hard parts"); a split-precision tcgen05 trailing update is future work.
#include "krylov_grid.cuh"
#include "krylov_grid_api.cuh"

namespace lxb {

constexpr int kPB = 32;  // panel width

// v(i, c): entry of the unit-lower-trapezoidal V of the panel starting at global column/row j0
template <typename T>
__device__ __forceinline__ T vmask(T stored, int grow, int j0, int c) {
  const int d = j0 + c;
  return grow > d ? stored : (grow == d ? T(1) : T(0));
}

// ------------------------------------------------------------------ K1: panel ----
template <typename T>
__global__ void __launch_bounds__(kGridThreads)
    qr_panel_kernel(T* __restrict__ a, T* __restrict__ taus, T* __restrict__ Tout, T* __restrict__ part,
                    T* __restrict__ gpart, int m, int n, int j0, int nbw, int in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* red = reinterpret_cast<T*>(smem_raw);   // 96 + kGridMaxK
  T* vals = red + 96 + kGridMaxK;            // 64: [0,32) Gram row, [32,64) pivot row
  T* fc = vals + 64;                         // 32
  T* Ts = fc + 32;                           // 32 x 32 T
  T* Gs = Ts + kPB * kPB;                    // 32 x 32 Gram of V
  T* wsum = Gs + kPB * kPB;                  // 8 x 32 cross-warp scratch
  T* Psm = wsum + 8 * 32;                    // rows_cta x 33 (shared-memory mode)
  GridTeam<T> team(part, red);
  const int tid = team.tid, nt = team.nt, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  int lo, hi;
  team.slice(m - j0, lo, hi);  // local row indices within [j0, m)
  const int nrows = hi - lo;
  // The CTA's rows of the panel live in shared memory when they fit, otherwise they are worked on
  // in place in global memory (the 32-column panel of a 262144-row matrix is 33.5 MB: L2-resident).
  T* P = in_smem ? Psm : a + (size_t)(j0 + lo) * n + j0;
  const size_t ldp = in_smem ? 33 : (size_t)n;
  const bool cok = lane < nbw;  // column guard (last, partial panel)
  if (in_smem) {
    for (int idx = tid; idx < nrows * kPB; idx += nt) {
      const int r = idx / kPB, c = idx % kPB;
      Psm[r * 33 + c] = c < nbw ? a[(size_t)(j0 + lo + r) * n + j0 + c] : T(0);
    }
  }
  __syncthreads();
  for (int jj = 0; jj < nbw; ++jj) {
    const int jrow = jj;  // local (panel) row index of the diagonal element: global row j0 + jj
    // Gram row over own rows strictly below the diagonal + the pivot row itself
    // 8 rows in flight per warp: the panel is L2-resident in global mode, so the loads must overlap
    T acc = T(0);
    for (int r0 = warp; r0 < nrows; r0 += 8 * nw) {
      T pj[8], pc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = r0 + q * nw;
        const bool ok = r < nrows && lo + r > jrow && cok;
        pj[q] = ok ? P[r * ldp + jj] : T(0);
        pc[q] = ok ? P[r * ldp + lane] : T(0);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fma_(pj[q], pc[q], acc);
    }
    wsum[warp * 32 + lane] = acc;
    __syncthreads();
    if (warp == 0) {
      T s = T(0);
      for (int w = 0; w < nw; ++w) s += wsum[w * 32 + lane];
      vals[lane] = s;
      const bool owner = jrow >= lo && jrow < hi;
      vals[32 + lane] = (owner && cok) ? P[(jrow - lo) * ldp + lane] : T(0);
    }
    __syncthreads();
    team.reduce_dyn(vals, 64);
    // larfg
    const T alpha = vals[32 + jj], ssq = vals[jj];
    T tau = T(0), beta = alpha, scal = T(1);
    if (ssq != T(0)) {
      const T nrm = sqrt_(alpha * alpha + ssq);
      beta = alpha >= T(0) ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scal = T(1) / (alpha - beta);
    }
    if (tid < 32) fc[tid] = tid > jj && tid < nbw ? tau * (vals[32 + tid] + scal * vals[tid]) : T(0);
    __syncthreads();
    // v = scal * a[:, jj] below the diagonal; trailing panel columns -= f_c v
    const T fcl = fc[lane];
    for (int r0 = warp; r0 < nrows; r0 += 8 * nw) {
      T pj[8], pc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = r0 + q * nw;
        const bool ok = r < nrows && lo + r >= jrow && cok;
        pj[q] = ok ? P[r * ldp + jj] : T(0);
        pc[q] = ok ? P[r * ldp + lane] : T(0);
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = r0 + q * nw;
        if (r >= nrows || !cok) continue;
        const int lr = lo + r;
        if (lr > jrow) {
          const T v = pj[q] * scal;
          if (lane == jj) P[r * ldp + jj] = v;
          else if (lane > jj) P[r * ldp + lane] = fma_(-fcl, v, pc[q]);
        } else if (lr == jrow) {
          if (lane == jj) P[r * ldp + jj] = beta;
          else if (lane > jj) P[r * ldp + lane] = pc[q] - fcl;
        }
      }
    }
    if (team.bid == 0 && tid == 0) taus[j0 + jj] = tau;
    if (tid == 0) Ts[jj * kPB + jj] = tau;  // diagonal of T
    __syncthreads();
  }
  // write the factored panel back
  if (in_smem) {
    for (int idx = tid; idx < nrows * kPB; idx += nt) {
      const int r = idx / kPB, c = idx % kPB;
      if (c < nbw) a[(size_t)(j0 + lo + r) * n + j0 + c] = Psm[r * 33 + c];
    }
  }
  // Gram of V (unit lower trapezoidal): G[c1][c2] = sum_i v(i,c1) v(i,c2), per-CTA partial -> global
  // warps split the rows (each row is read once as a coalesced 128-byte line and kept in a register
  // per lane); lane = column c2; the 32 columns c1 are broadcast with shuffles
  {
    T g[kPB];
#pragma unroll
    for (int c1 = 0; c1 < kPB; ++c1) g[c1] = T(0);
    for (int r = warp; r < nrows; r += nw) {
      const int gr = j0 + lo + r;
      const T mine = cok ? vmask<T>(P[r * ldp + lane], gr, j0, lane) : T(0);
#pragma unroll
      for (int c1 = 0; c1 < kPB; ++c1) g[c1] = fma_(__shfl_sync(kFull, mine, c1), mine, g[c1]);
    }
    // cross-warp sum through shared memory (Gs reused as scratch: 8 warps x 32 x 32 does not fit,
    // so accumulate warp by warp)
    for (int e = tid; e < kPB * kPB; e += nt) Gs[e] = T(0);
    __syncthreads();
    for (int w = 0; w < nw; ++w) {
      if (warp == w) {
#pragma unroll
        for (int c1 = 0; c1 < kPB; ++c1) Gs[c1 * kPB + lane] += g[c1];
      }
      __syncthreads();
    }
    for (int e = tid; e < kPB * kPB; e += nt) gpart[(size_t)team.bid * (kPB * kPB) + e] = Gs[e];
  }
  __threadfence();
  team.sync();
  for (int e = tid; e < kPB * kPB; e += nt) {
    T s = T(0);
    for (int b = 0; b < team.nb; ++b) s += __ldcg(gpart + (size_t)b * (kPB * kPB) + e);
    Gs[e] = s;
  }
  __syncthreads();
  // larft (forward, columnwise): T[0:j, j] = -tau_j * T[0:j, 0:j] * G[0:j, j]
  for (int j = 1; j < nbw; ++j) {
    const T tj = Ts[j * kPB + j];
    if (tid < j) {
      T s = T(0);
      for (int q = tid; q < j; ++q) s = fma_(Ts[tid * kPB + q], Gs[q * kPB + j], s);  // T upper: rows tid, cols q >= tid
      fc[tid] = -tj * s;
    }
    __syncthreads();
    if (tid < j) Ts[tid * kPB + j] = fc[tid];
    __syncthreads();
  }
  if (team.bid == 0) {
    for (int e = tid; e < kPB * kPB; e += nt) {
      const int r = e / kPB, c = e % kPB;
      Tout[e] = (c >= r && c < nbw && r < nbw) ? Ts[e] : T(0);
    }
  }
}

// ------------------------------------------------------------- K2: W partial ----
// Wp[g][k][c] = sum_{i in row group g} v(i,k) * A[i][cbase + c]; CTA tile = 32 k x 256 columns,
// thread tile 4 k x 8 columns (32 FMA per 3 LDS.128); the next 32-row chunk is prefetched from
// global memory into registers while the current one is consumed from shared memory.
constexpr int kW2Tile = 256;
template <typename T>
__global__ void __launch_bounds__(256)
    qr_wpartial_kernel(const T* __restrict__ a, T* __restrict__ Wp, int m, int n, int j0, int ncols,
                       int ngroups) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T(*Vs)[kPB] = reinterpret_cast<T(*)[kPB]>(smem_raw);                              // [32][32]
  T(*As)[kW2Tile] = reinterpret_cast<T(*)[kW2Tile]>(smem_raw + 32 * kPB * sizeof(T));  // [32][256]
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, grp = blockIdx.y;
  const int cbase = j0 + kPB + tile * kW2Tile;
  const int cw = min(kW2Tile, j0 + kPB + ncols - cbase);
  const int rows_total = m - j0;
  const int per = (((rows_total + ngroups - 1) / ngroups) + 31) & ~31;
  const int r0 = j0 + grp * per, r1 = min(m, r0 + per);
  const int kq = tid >> 5, cq = tid & 31;  // 8 groups of 4 k's, 32 groups of 8 columns
  T acc[4][8];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 8; ++y) acc[x][y] = T(0);
  // staging assignment: V chunk 32x32 -> 4 elements per thread; A chunk 32x256 -> 32 per thread
  T pv[4], pa[32];
  auto fetch = [&](int rb) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * 256, r = idx / kPB, c = idx % kPB, gr = rb + r;
      pv[q] = gr < r1 ? vmask<T>(a[(size_t)gr * n + j0 + c], gr, j0, c) : T(0);
    }
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int idx = tid + q * 256, r = idx / kW2Tile, c = idx % kW2Tile, gr = rb + r;
      pa[q] = (gr < r1 && c < cw) ? a[(size_t)gr * n + cbase + c] : T(0);
    }
  };
  if (r0 < r1) fetch(r0);
  for (int rb = r0; rb < r1; rb += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * 256;
      Vs[idx / kPB][idx % kPB] = pv[q];
    }
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int idx = tid + q * 256;
      As[idx / kW2Tile][idx % kW2Tile] = pa[q];
    }
    __syncthreads();
    if (rb + 32 < r1) fetch(rb + 32);
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      T v[4], x[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = Vs[r][4 * kq + e];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = As[r][8 * cq + e];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[p][q] = fma_(v[p], x[q], acc[p][q]);
    }
    __syncthreads();
  }
  T* out = Wp + ((size_t)grp * kPB) * ncols;
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = tile * kW2Tile + 8 * cq + q;
      if (c < ncols) out[(size_t)(4 * kq + p) * ncols + c] = acc[p][q];
    }
}

// -------------------------------------------------------------- K3: W finish ----
// W2[k][c] = sum_k' T[k'][k] * (sum_g Wp[g][k'][c])
template <typename T>
__global__ void __launch_bounds__(256)
    qr_wfinish_kernel(const T* __restrict__ Wp, const T* __restrict__ Tm, T* __restrict__ W2, int ncols,
                      int ngroups) {
  __shared__ T Ts[kPB * kPB];
  for (int e = threadIdx.x; e < kPB * kPB; e += blockDim.x) Ts[e] = Tm[e];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  T w[kPB];
#pragma unroll
  for (int k = 0; k < kPB; ++k) {
    T s = T(0);
    for (int g = 0; g < ngroups; ++g) s += Wp[((size_t)g * kPB + k) * ncols + c];
    w[k] = s;
  }
#pragma unroll
  for (int k = 0; k < kPB; ++k) {
    T s = T(0);
#pragma unroll
    for (int kp = 0; kp <= k; ++kp) s = fma_(Ts[kp * kPB + k], w[kp], s);  // T upper triangular
    W2[(size_t)k * ncols + c] = s;
  }
}

// ------------------------------------------------------------------ K4: update ----
// A[i][cbase + c] -= sum_k v(i,k) W2[k][c]; CTA tile 64 rows x 128 columns, thread 4 x 8.
template <typename T>
__global__ void __launch_bounds__(256)
    qr_update_kernel(T* __restrict__ a, const T* __restrict__ W2, int m, int n, int j0, int ncols) {
  __shared__ __align__(16) T Vt[kPB][64];
  __shared__ __align__(16) T Ws[kPB][128];
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, rblk = blockIdx.y;
  const int cbase = j0 + kPB + tile * 128;
  const int rb = j0 + rblk * 64;
  for (int idx = tid; idx < 64 * kPB; idx += 256) {
    const int r = idx / kPB, c = idx % kPB, gr = rb + r;
    Vt[c][r] = gr < m ? vmask<T>(a[(size_t)gr * n + j0 + c], gr, j0, c) : T(0);
  }
  for (int idx = tid; idx < kPB * 128; idx += 256) {
    const int k = idx / 128, c = idx % 128;
    const int gc = tile * 128 + c;
    Ws[k][c] = gc < ncols ? W2[(size_t)k * ncols + gc] : T(0);
  }
  __syncthreads();
  const int ri = tid >> 4, ci = tid & 15;  // 16 row groups of 4, 16 column groups of 8
  T acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[p][q] = T(0);
#pragma unroll 8
  for (int k = 0; k < kPB; ++k) {
    T v[4], w[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) v[p] = Vt[k][4 * ri + p];
#pragma unroll
    for (int q = 0; q < 8; ++q) w[q] = Ws[k][8 * ci + q];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[p][q] = fma_(v[p], w[q], acc[p][q]);
  }
  using VT = typename V16K<T>::type;
  constexpr int V = 16 / sizeof(T);
  const bool vec = (n % V == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                   (tile * 128 + 8 * ci + 8 <= ncols);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int gr = rb + 4 * ri + p;
    if (gr >= m) continue;
    T* rowp = a + (size_t)gr * n + cbase + 8 * ci;
    if (vec) {
#pragma unroll
      for (int h = 0; h < 8 / V; ++h) {
        VT val = reinterpret_cast<VT*>(rowp)[h];
        T* pv = reinterpret_cast<T*>(&val);
#pragma unroll
        for (int e = 0; e < V; ++e) pv[e] = pv[e] - acc[p][h * V + e];
        reinterpret_cast<VT*>(rowp)[h] = val;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int gc = tile * 128 + 8 * ci + q;
        if (gc < ncols) rowp[q] = rowp[q] - acc[p][q];
      }
    }
  }
}

// --------------------------------------------------- apply Q^T to one vector ----
// y <- H_n ... H_2 H_1 y, panel by panel with the panel rows in shared memory (cooperative).
template <typename T>
__global__ void __launch_bounds__(kGridThreads)
    qr_apply_qt_kernel(const T* __restrict__ a, const T* __restrict__ taus, T* __restrict__ y,
                       T* __restrict__ part, int m, int n, int in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* red = reinterpret_cast<T*>(smem_raw);
  T* vals = red + 96 + kGridMaxK;  // 1
  T* P = vals + 8;                 // rows_cta x 33
  GridTeam<T> team(part, red);
  const int tid = team.tid, nt = team.nt;
  for (int j0 = 0; j0 < n; j0 += kPB) {
    const int nbw = min(kPB, n - j0);
    int lo, hi;
    team.slice(m - j0, lo, hi);
    const int nrows = hi - lo;
    if (in_smem) {
      for (int idx = tid; idx < nrows * kPB; idx += nt) {
        const int r = idx / kPB, c = idx % kPB;
        P[r * 33 + c] = c < nbw ? vmask<T>(a[(size_t)(j0 + lo + r) * n + j0 + c], j0 + lo + r, j0, c) : T(0);
      }
    }
    __syncthreads();
    for (int jj = 0; jj < nbw; ++jj) {
      T d[1] = {T(0)};
      for (int r = tid; r < nrows; r += nt) {
        const int gr = j0 + lo + r;
        const T v = in_smem ? P[r * 33 + jj] : vmask<T>(a[(size_t)gr * n + j0 + jj], gr, j0, jj);
        d[0] = fma_(v, y[gr], d[0]);
      }
      team.template reduce<1, 0>(d, nullptr);
      const T f = taus[j0 + jj] * d[0];
      for (int r = tid; r < nrows; r += nt) {
        const int gr = j0 + lo + r;
        const T v = in_smem ? P[r * 33 + jj] : vmask<T>(a[(size_t)gr * n + j0 + jj], gr, j0, jj);
        y[gr] = fma_(-f, v, y[gr]);
      }
      __syncthreads();
    }
    team.sync();  // rows are re-partitioned for the next panel
  }
}

// x = R^{-1} y[0:n] for the n x n upper-triangular R stored in `a` (row-major, ld n): one CTA,
// column-oriented with the vector in shared memory.
template <typename T>
__global__ void __launch_bounds__(1024) qr_rsolve_kernel(const T* __restrict__ a, const T* __restrict__ y,
                                                         T* __restrict__ x, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n; i += nt) s[i] = y[i];
  __syncthreads();
  for (int k = n - 1; k >= 0; --k) {
    if (tid == 0) s[k] = s[k] / a[(size_t)k * n + k];
    __syncthreads();
    const T xk = s[k];
    for (int i = tid; i < k; i += nt) s[i] = fma_(-a[(size_t)i * n + k], xk, s[i]);
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) x[i] = s[i];
}

// ------------------------------------------------------------------ host side ----
template <typename T>
struct QrLargePlan {
  int nb, rows_cta;
  size_t smem_panel, smem_apply, ws_bytes, wp_off, w2_off, t_off, gpart_off, y_off;
  int ngroups, in_smem;
  bool ok;
};

template <typename T>
QrLargePlan<T> qr_large_plan(int m, int n) {
  QrLargePlan<T> pl{};
  pl.nb = grid_blocks();
  const int per = (((m + pl.nb - 1) / pl.nb) + 3) & ~3;
  pl.rows_cta = per;
  const size_t fixed_panel = (96 + kGridMaxK + 64 + 32 + 2 * kPB * kPB + 8 * 32) * sizeof(T);
  const size_t fixed_apply = (96 + kGridMaxK + 8) * sizeof(T);
  const size_t rows_bytes = (size_t)per * 33 * sizeof(T);
  // both cooperative kernels run 2 CTAs per SM; the row block goes to shared memory only if it fits
  pl.in_smem = (fixed_panel + rows_bytes) * kGridCtasPerSm <= 220 * 1024;
  pl.smem_panel = fixed_panel + (pl.in_smem ? rows_bytes : 0);
  pl.smem_apply = fixed_apply + (pl.in_smem ? rows_bytes : 0);
  pl.ngroups = 32;
  size_t off = grid_part_elems();
  pl.gpart_off = off; off += (size_t)pl.nb * kPB * kPB;
  pl.t_off = off; off += kPB * kPB;
  pl.wp_off = off; off += (size_t)pl.ngroups * kPB * pad4(n);
  pl.w2_off = off; off += (size_t)kPB * pad4(n);
  pl.y_off = off; off += pad4(m);
  pl.ws_bytes = off * sizeof(T);
  pl.ok = true;
  return pl;
}

template <typename T>
int qr_large_factor(const T* A, T* a, T* taus, int m, int n, void* ws, size_t ws_bytes, cudaStream_t st) {
  const QrLargePlan<T> pl = qr_large_plan<T>(m, n);
  if (!pl.ok) return LXB_E_UNSUPPORTED;
  if (!ws || ws_bytes < pl.ws_bytes) return LXB_E_WORKSPACE;
  T* w = reinterpret_cast<T*>(ws);
  LXB_CUDA_TRY(cudaMemcpyAsync(a, A, (size_t)m * n * sizeof(T), cudaMemcpyDeviceToDevice, st));
  count_launch();
  auto pk = qr_panel_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_panel));
  int occ = 0;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pk, kGridThreads, pl.smem_panel));
  if (occ * kNumSMs < pl.nb) return LXB_E_UNSUPPORTED;
  T* part = w;
  T* gpart = w + pl.gpart_off;
  T* Tm = w + pl.t_off;
  T* Wp = w + pl.wp_off;
  T* W2 = w + pl.w2_off;
  for (int j0 = 0; j0 < n; j0 += kPB) {
    int nbw = n - j0 < kPB ? n - j0 : kPB;
    int mm = m, nn = n, jj0 = j0, ism = pl.in_smem;
    void* args[] = {&a, &taus, &Tm, &part, &gpart, &mm, &nn, &jj0, &nbw, &ism};
    LXB_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)pk, dim3(pl.nb), dim3(kGridThreads), args,
                                             pl.smem_panel, st));
    count_launch();
    const int ncols = n - j0 - kPB;
    if (ncols <= 0) break;
    const int tiles = (ncols + 127) / 128;
    const int wtiles = (ncols + kW2Tile - 1) / kW2Tile;
    int ngroups = pl.ngroups;
    const size_t w_smem = (size_t)(32 * kPB + 32 * kW2Tile) * sizeof(T);
    LXB_CUDA_TRY(cudaFuncSetAttribute(qr_wpartial_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w_smem));
    qr_wpartial_kernel<T><<<dim3(wtiles, ngroups), 256, w_smem, st>>>(a, Wp, m, n, j0, ncols, ngroups);
    LXB_CUDA_CHECK_LAUNCH();
    qr_wfinish_kernel<T><<<(ncols + 255) / 256, 256, 0, st>>>(Wp, Tm, W2, ncols, ngroups);
    LXB_CUDA_CHECK_LAUNCH();
    const int rblocks = (m - j0 + 63) / 64;
    qr_update_kernel<T><<<dim3(tiles, rblocks), 256, 0, st>>>(a, W2, m, n, j0, ncols);
    LXB_CUDA_CHECK_LAUNCH();
  }
  return 0;
}

// least squares with the large factors: x = R^{-1} (Q^T b)[:n]
template <typename T>
int qr_large_solve(const T* a, const T* taus, const T* b, T* x, int m, int n, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  const QrLargePlan<T> pl = qr_large_plan<T>(m, n);
  if (!pl.ok) return LXB_E_UNSUPPORTED;
  if (!ws || ws_bytes < pl.ws_bytes) return LXB_E_WORKSPACE;
  T* w = reinterpret_cast<T*>(ws);
  T* part = w;
  T* y = w + pl.y_off;
  LXB_CUDA_TRY(cudaMemcpyAsync(y, b, (size_t)m * sizeof(T), cudaMemcpyDeviceToDevice, st));
  count_launch();
  auto ak = qr_apply_qt_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(ak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_apply));
  int mm = m, nn = n, ism = pl.in_smem;
  void* args[] = {&a, &taus, &y, &part, &mm, &nn, &ism};
  LXB_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)ak, dim3(pl.nb), dim3(kGridThreads), args,
                                           pl.smem_apply, st));
  count_launch();
  const size_t smem = (size_t)n * sizeof(T);
  if (smem > 200 * 1024) return LXB_E_UNSUPPORTED;
  LXB_CUDA_TRY(cudaFuncSetAttribute(qr_rsolve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qr_rsolve_kernel<T><<<1, 1024, smem, st>>>(a, y, x, n);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
size_t qr_large_ws_bytes(int m, int n) { return qr_large_plan<T>(m, n).ws_bytes; }

#define LXB_INST_QRL(T)                                                                             \
  template int qr_large_factor<T>(const T*, T*, T*, int, int, void*, size_t, cudaStream_t);         \
  template int qr_large_solve<T>(const T*, const T*, const T*, T*, int, int, void*, size_t, cudaStream_t); \
  template size_t qr_large_ws_bytes<T>(int, int);
LXB_INST_QRL(float)
LXB_INST_QRL(double)

}  // namespace lxb

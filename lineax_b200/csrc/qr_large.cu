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
// The 1e-5 parity budget rules out plain TF32 tensor-core tiles (SURVEY.md section 7 "hard parts"):
// the fp32 updates run on tcgen05 with the 3xTF32 split (qr_update_tc_kernel for the 32-wide ones).
// fp32, aligned, n > 128 uses TWO-LEVEL blocking: four panels per 128-column outer block (K2-K4 only
// on the <= 96 columns inside the block, partials pre-reduced by qr_wpresum_kernel), then once per
// outer block, K = 128 on the tensor cores:
//     qr_wbig_tc_kernel<true> + qr_gsum + qr_tbig : T (128 x 128) from the Gram matrix of V
//     qr_vsplit_kernel      : V -> pre-split hi/lo operand images (K-major and MN-major)
//     qr_wbig_ws_kernel     : Wp[g] = V^T A2, warp specialised, MN-major operands
//     qr_wfinish128_kernel  : Y = T^T sum_g Wp[g]
//     qr_update128_ws_kernel: A2 -= V Y, warp specialised, bulk-copied V images
// Everything else is fp32/fp64 SIMT FMA.
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
// Cooperative kernel, ONE 512-thread CTA per SM.  A CTA keeps its rows of the 32-column panel ON
// CHIP for the whole factorisation: the first RR * 16 rows in registers (warp w holds rows w,
// w+16, ...; lane = column, so a column broadcast is one shuffle) and the rest in shared memory
// (in place in global memory, L2 resident, only when even that does not fit: XS = false).
// Per column there is ONE sweep over the rows -- it applies reflector jj and accumulates the Gram
// row of column jj+1 on the fly -- and ONE grid barrier: every CTA publishes its 32 partial Gram
// sums (the pivot row's owner also the pivot row), and after the barrier every CTA sums the nb
// partials itself, all loads in flight at once (measured on B200: the partial read is the
// critical path of the column, 15 us when its L2 round trips serialise, < 2 us like this).
constexpr int kPanelThreads = 512;
constexpr int kPanelWarps = kPanelThreads / 32;
template <typename T>
struct PanelCfg {
  static constexpr int RR = sizeof(T) == 4 ? 40 : 20;  // register rows per warp
  static constexpr int fixed_elems = 64 + 32 + 2 * kPB * kPB + kPanelWarps * 32 + 32;
};

template <typename T, bool XS>
__global__ void __launch_bounds__(kPanelThreads, 1)
    qr_panel_kernel(T* __restrict__ a, T* __restrict__ taus, T* __restrict__ Tout, T* __restrict__ part,
                    T* __restrict__ gpart, T* __restrict__ gfull, int m, int n, int j0, int nbw) {
  constexpr int RR = PanelCfg<T>::RR;
  constexpr int nw = kPanelWarps;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* vals = reinterpret_cast<T*>(smem_raw);  // 64: [0,32) Gram row, [32,64) pivot row
  T* fc = vals + 64;                         // 32
  T* Ts = fc + 32;                           // 32 x 32 T
  T* Gs = Ts + kPB * kPB;                    // 32 x 32 Gram of V
  T* wsum = Gs + kPB * kPB;                  // 16 x 32 cross-warp scratch
  T* piv = wsum + nw * 32;                   // 32: pivot row of the next column
  T* Psm = piv + 32;                         // extra rows x 33 (XS)
  cgx::grid_group grid = cgx::this_grid();
  const int nb = gridDim.x, bid = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = ((((m - j0) + nb - 1) / nb) + 3) & ~3;
  const int lo = bid * per < m - j0 ? bid * per : m - j0;  // local row indices within [j0, m)
  const int hi = lo + per < m - j0 ? lo + per : m - j0;
  const int nrows = hi - lo;
  const int nreg = nrows < RR * nw ? nrows : RR * nw;  // local rows [0, nreg): registers
  const int nx = nrows - nreg;                         // local rows [nreg, nrows): X
  T* Xg = a + (size_t)(j0 + lo + nreg) * n + j0;
  auto X = [&](int r, int c) -> T& { return XS ? Psm[r * 33 + c] : Xg[(size_t)r * n + c]; };
  const bool cok = lane < nbw;  // column guard (last, partial panel)
  // global scratch: 2 x (32 x nb partial sums + 32 pivot-row values), double buffered by column
  const size_t pstride = (size_t)kPB * nb + kPB;
#ifdef LXB_QR_PROF
  if (blockIdx.x == 0 && threadIdx.x == 0) g_prof[15] = prof_now();
#endif
  T reg[RR];
#pragma unroll
  for (int q = 0; q < RR; ++q) {
    const int r = q * nw + warp;
    reg[q] = (r < nreg && cok) ? a[(size_t)(j0 + lo + r) * n + j0 + lane] : T(0);
  }
  if (XS) {
    for (int idx = tid; idx < nx * kPB; idx += kPanelThreads) {
      const int r = idx / kPB, c = idx % kPB;
      Psm[r * 33 + c] = c < nbw ? Xg[(size_t)r * n + c] : T(0);
    }
  }
  __syncthreads();
  // Loop trip counts below are written in terms of block-uniform values only, so that the compiler
  // can prove the shuffles convergent (otherwise every one of them is wrapped in a WARPSYNC).
  const int xiters = (nx + 8 * nw - 1) / (8 * nw);
  // Gram row of column 0 over the rows strictly below the diagonal (+ its pivot row); later
  // columns get theirs from the update sweep of the previous column.
  T carry = T(0);
#pragma unroll
  for (int q = 0; q < RR; ++q) {
    const int r = q * nw + warp, lr = lo + r;
    const T pj = __shfl_sync(kFull, reg[q], 0);
    carry = (r < nreg && lr > 0) ? fma_(pj, reg[q], carry) : carry;
    if (r < nreg && lr == 0) piv[lane] = reg[q];
  }
  for (int it = 0; it < xiters; ++it) {
    const int r0 = warp + it * 8 * nw;
    T pj[8], pc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = r0 + q * nw;
      const bool ok = r < nx && cok;
      pj[q] = ok ? X(r, 0) : T(0);
      pc[q] = ok ? X(r, lane) : T(0);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) carry = fma_(pj[q], pc[q], carry);
  }
  for (int jj = 0; jj < nbw; ++jj) {
    const int jrow = jj;  // local (panel) row index of the diagonal element: global row j0 + jj
    const T acc = carry;
    wsum[warp * 32 + lane] = acc;
    __syncthreads();
    LXB_PROF(0);
    T* buf = part + (size_t)(jj & 1) * pstride;
    if (warp == 0) {
      T s = T(0);
#pragma unroll
      for (int w = 0; w < nw; ++w) s += wsum[w * 32 + lane];
      buf[(size_t)lane * nb + bid] = s;
      if (jrow >= lo && jrow < hi) buf[(size_t)kPB * nb + lane] = cok ? piv[lane] : T(0);
      __threadfence();
    }
    LXB_PROF(1);
    grid.sync();
    LXB_PROF(2);
    {
      // warp w sums entries 2w and 2w+1 over the nb partials, lanes striding the CTAs
      const T* b0 = buf + (size_t)(2 * warp) * nb;
      const T* b1 = b0 + nb;
      T s0 = T(0), s1 = T(0);
      for (int i0 = lane; i0 < nb; i0 += 32 * 5) {
        T t0[5], t1[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const int i = i0 + 32 * u;
          t0[u] = i < nb ? __ldcg(b0 + i) : T(0);
          t1[u] = i < nb ? __ldcg(b1 + i) : T(0);
        }
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          s0 += t0[u];
          s1 += t1[u];
        }
      }
      T pv = T(0);
      if (warp == 0) pv = __ldcg(buf + (size_t)kPB * nb + lane);
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) {
        vals[2 * warp] = s0;
        vals[2 * warp + 1] = s1;
      }
      if (warp == 0) vals[32 + lane] = pv;
    }
    __syncthreads();
    LXB_PROF(3);
    // larfg
    const T alpha = vals[32 + jj], ssq = vals[jj];
    T tau = T(0), beta = alpha, scal = T(1);
    if (ssq != T(0)) {
      const T nrm = sqrt_(alpha * alpha + ssq);
      beta = alpha >= T(0) ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scal = T(1) / (alpha - beta);
    }
    // f_c = tau * (pivot-row entry + v^T a_c) for the panel columns right of jj
    const T fcl = (lane > jj && cok) ? tau * (vals[32 + lane] + scal * vals[lane]) : T(0);
    LXB_PROF(4);
    // v = scal * a[:, jj] below the diagonal; trailing panel columns -= f_c v; and, in the same
    // sweep, the Gram row of column jj + 1 from the updated values.  Branch-free on purpose.
    const int jn = jj + 1;
    const bool more = jn < nbw;
    const bool right = lane > jj;
    carry = T(0);
#pragma unroll
    for (int q = 0; q < RR; ++q) {
      const int r = q * nw + warp, lr = lo + r;
      const bool live = r < nreg && cok;
      const T old = reg[q];
      const T v = __shfl_sync(kFull, old, jj) * scal;
      const T below = right ? fma_(-fcl, v, old) : v;  // rows under the diagonal
      const T ondiag = right ? old - fcl : beta;       // the pivot row
      const T cand = lr > jrow ? below : ondiag;
      const T nv = (live && lr >= jrow && lane >= jj) ? cand : old;
      reg[q] = nv;
      const T pn = __shfl_sync(kFull, nv, jn & 31);
      carry = (more && live && lr > jrow + 1) ? fma_(pn, nv, carry) : carry;
      if (more && r < nreg && lr == jrow + 1) piv[lane] = cok ? nv : T(0);
    }
    // extra rows (always strictly below the panel's diagonal block: nreg >= 32 whenever nx > 0)
    for (int it = 0; it < xiters; ++it) {
      const int r0 = warp + it * 8 * nw;
      T pc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = r0 + q * nw;
        pc[q] = (r < nx && cok) ? X(r, lane) : T(0);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = r0 + q * nw;
        const bool live = r < nx && cok;
        const T v = __shfl_sync(kFull, pc[q], jj) * scal;
        const T nv = lane >= jj ? (right ? fma_(-fcl, v, pc[q]) : v) : pc[q];
        if (live && lane >= jj) X(r, lane) = nv;
        const T pn = __shfl_sync(kFull, nv, jn & 31);
        carry = (more && live) ? fma_(pn, nv, carry) : carry;
      }
    }
    if (bid == 0 && tid == 0) taus[j0 + jj] = tau;
    if (tid == 0) Ts[jj * kPB + jj] = tau;  // diagonal of T
    LXB_PROF(5);
  }
  __syncthreads();
  // write the factored panel back
#pragma unroll
  for (int q = 0; q < RR; ++q) {
    const int r = q * nw + warp;
    if (r < nreg && cok) a[(size_t)(j0 + lo + r) * n + j0 + lane] = reg[q];
  }
  if (XS) {
    for (int idx = tid; idx < nx * kPB; idx += kPanelThreads) {
      const int r = idx / kPB, c = idx % kPB;
      if (c < nbw) Xg[(size_t)r * n + c] = Psm[r * 33 + c];
    }
  }
  // Gram of V (unit lower trapezoidal): G[c1][c2] = sum_i v(i,c1) v(i,c2), per-CTA partial -> global.
  // lane = column c2; the 32 columns c1 are broadcast with shuffles
  {
    T g[kPB];
#pragma unroll
    for (int c1 = 0; c1 < kPB; ++c1) g[c1] = T(0);
#pragma unroll
    for (int q = 0; q < RR; ++q) {
      const int r = q * nw + warp;
      const T mine = (r < nreg && cok) ? vmask<T>(reg[q], j0 + lo + r, j0, lane) : T(0);
#pragma unroll
      for (int c1 = 0; c1 < kPB; ++c1) g[c1] = fma_(__shfl_sync(kFull, mine, c1), mine, g[c1]);
    }
    for (int it = 0; it < (nx + nw - 1) / nw; ++it) {
      const int r = warp + it * nw;
      const T mine = (r < nx && cok) ? X(r, lane) : T(0);
#pragma unroll
      for (int c1 = 0; c1 < kPB; ++c1) g[c1] = fma_(__shfl_sync(kFull, mine, c1), mine, g[c1]);
    }
    // cross-warp sum through shared memory, warp by warp (fixed order)
    for (int e = tid; e < kPB * kPB; e += kPanelThreads) Gs[e] = T(0);
    __syncthreads();
    for (int w = 0; w < nw; ++w) {
      if (warp == w) {
#pragma unroll
        for (int c1 = 0; c1 < kPB; ++c1) Gs[c1 * kPB + lane] += g[c1];
      }
      __syncthreads();
    }
    for (int e = tid; e < kPB * kPB; e += kPanelThreads) gpart[(size_t)bid * (kPB * kPB) + e] = Gs[e];
  }
  LXB_PROF(6);
  __threadfence();
  grid.sync();
  LXB_PROF(7);
  // two-level sum of the per-CTA Gram partials: CTA b < 32 reduces row b, CTA 0 then builds T
  if (bid < kPB) {
    const int e0 = bid * kPB + warp * 2;  // 2 entries per warp
    T s[2] = {T(0), T(0)};
    for (int b = lane; b < nb; b += 32) {
#pragma unroll
      for (int u = 0; u < 2; ++u) s[u] += __ldcg(gpart + (size_t)b * (kPB * kPB) + e0 + u);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const T t = warp_sum(s[u]);
      if (lane == 0) gfull[e0 + u] = t;
    }
  }
  __threadfence();
  grid.sync();
  LXB_PROF(8);
  if (bid != 0) return;
  for (int e = tid; e < kPB * kPB; e += kPanelThreads) Gs[e] = __ldcg(gfull + e);
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
  for (int e = tid; e < kPB * kPB; e += kPanelThreads) {
    const int r = e / kPB, c = e % kPB;
    Tout[e] = (c >= r && c < nbw && r < nbw) ? Ts[e] : T(0);
  }
  LXB_PROF(9);
}

// ------------------------------------------------------------- K2: W partial ----
// Wp[g][k][c] = sum_{i in row group g} v(i,k) * A[i][cbase + c].  CTA tile = 32 k x 256 columns,
// 4 warps; warp w owns k = 8w..8w+7, lane l owns columns 4l..4l+3 and 128+4l..128+4l+3, so one
// row step is 64 FMA per 4 LDS.128 (two of them warp-wide broadcasts, two conflict-free).
// 16-row chunks of V and A2 travel global -> shared with cp.async (zero-filled past the group /
// tile edge) through a 3-stage ring, so no staging registers and no STS in the loop.
constexpr int kW2Tile = 256;
constexpr int kW2Rows = 16;
constexpr int kW2Stages = 3;
constexpr int kW2Threads = 128;
constexpr int kW2MaxGroups = 128;

__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gmem_src, int bytes,
                                               int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (bytes == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes));
  else if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes));
}

// 4 consecutive elements from 16-byte aligned shared memory
template <typename T>
__device__ __forceinline__ void lds4(const T* p, T* out) {
  using VT = typename V16K<T>::type;
  constexpr int V = 16 / sizeof(T);
#pragma unroll
  for (int h = 0; h < 4 / V; ++h) {
    const VT t = reinterpret_cast<const VT*>(p)[h];
    const T* pt = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int e = 0; e < V; ++e) out[h * V + e] = pt[e];
  }
}

template <typename T>
__global__ void __launch_bounds__(kW2Threads, sizeof(T) == 4 ? 4 : 2)
    qr_wpartial_kernel(const T* __restrict__ a, T* __restrict__ Wp, int m, int n, int j0, int ncols,
                       int ngroups) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kStageElems = kW2Rows * (kPB + kW2Tile);
  T* stage0 = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, grp = blockIdx.y;
  const int cbase = j0 + kPB + tile * kW2Tile;
  const int cw = min(kW2Tile, j0 + kPB + ncols - cbase);
  const int rows_total = m - j0;
  const int per = (((rows_total + ngroups - 1) / ngroups) + kW2Rows - 1) / kW2Rows * kW2Rows;
  const int r0 = j0 + grp * per, r1 = min(m, r0 + per);
  constexpr int V = 16 / sizeof(T);
  // 16-byte pieces need aligned row starts (j0 and cbase are multiples of 32 elements)
  const bool al = (n % V == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0);
  T acc[8][8];
#pragma unroll
  for (int x = 0; x < 8; ++x)
#pragma unroll
    for (int y = 0; y < 8; ++y) acc[x][y] = T(0);

  auto issue = [&](int st, int rb) {
    T* Vs = stage0 + (size_t)st * kStageElems;  // [16][32]
    T* As = Vs + kW2Rows * kPB;                 // [16][256]
    if (al) {
      constexpr int vp = kPB / V, ap = kW2Tile / V;  // pieces per row
#pragma unroll
      for (int q = 0; q < kW2Rows * vp / kW2Threads; ++q) {
        const int pc = tid + q * kW2Threads, r = pc / vp, c = (pc % vp) * V, gr = rb + r;
        const bool ok = gr < r1;
        cp_async_zfill(Vs + r * kPB + c, ok ? a + (size_t)gr * n + j0 + c : a, 16, ok ? 16 : 0);
      }
#pragma unroll
      for (int q = 0; q < kW2Rows * ap / kW2Threads; ++q) {
        const int pc = tid + q * kW2Threads, r = pc / ap, c = (pc % ap) * V, gr = rb + r;
        int nb = gr < r1 ? (cw - c) * (int)sizeof(T) : 0;
        nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
        cp_async_zfill(As + r * kW2Tile + c, nb ? a + (size_t)gr * n + cbase + c : a, 16, nb);
      }
    } else {
      for (int e = tid; e < kW2Rows * kPB; e += kW2Threads) {
        const int r = e / kPB, c = e % kPB, gr = rb + r;
        const bool ok = gr < r1;
        cp_async_zfill(Vs + e, ok ? a + (size_t)gr * n + j0 + c : a, (int)sizeof(T), ok ? (int)sizeof(T) : 0);
      }
      for (int e = tid; e < kW2Rows * kW2Tile; e += kW2Threads) {
        const int r = e / kW2Tile, c = e % kW2Tile, gr = rb + r;
        const bool ok = gr < r1 && c < cw;
        cp_async_zfill(As + e, ok ? a + (size_t)gr * n + cbase + c : a, (int)sizeof(T),
                       ok ? (int)sizeof(T) : 0);
      }
    }
  };

  const int nchunks = r1 > r0 ? (r1 - r0 + kW2Rows - 1) / kW2Rows : 0;
#pragma unroll
  for (int sidx = 0; sidx < kW2Stages - 1; ++sidx) {
    if (sidx < nchunks) issue(sidx, r0 + sidx * kW2Rows);
    cp_async_commit();
  }
  for (int ch = 0; ch < nchunks; ++ch) {
    cp_async_wait<kW2Stages - 2>();
    __syncthreads();
    const int nx = ch + kW2Stages - 1;
    if (nx < nchunks) issue(nx % kW2Stages, r0 + nx * kW2Rows);
    cp_async_commit();
    T* Vs = stage0 + (size_t)(ch % kW2Stages) * kStageElems;
    const T* As = Vs + kW2Rows * kPB;
    const int rb = r0 + ch * kW2Rows;
    if (rb < j0 + kPB) {  // rows crossing the panel's diagonal block: unit diagonal, zeros above
      for (int e = tid; e < kW2Rows * kPB; e += kW2Threads) {
        const int r = e / kPB, c = e % kPB, gr = rb + r;
        Vs[e] = gr < r1 ? vmask<T>(Vs[e], gr, j0, c) : T(0);
      }
      __syncthreads();
    }
#pragma unroll 2
    for (int r = 0; r < kW2Rows; ++r) {
      T v[8], x[8];
      lds4<T>(Vs + r * kPB + 8 * warp, v);
      lds4<T>(Vs + r * kPB + 8 * warp + 4, v + 4);
      lds4<T>(As + r * kW2Tile + 4 * lane, x);
      lds4<T>(As + r * kW2Tile + 128 + 4 * lane, x + 4);
#pragma unroll
      for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[p][q] = fma_(v[p], x[q], acc[p][q]);
    }
  }
  cp_async_wait<0>();
  T* out = Wp + ((size_t)grp * kPB) * ncols;
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = tile * kW2Tile + (q < 4 ? 4 * lane + q : 128 + 4 * lane + q - 4);
      if (c < ncols) out[(size_t)(8 * warp + p) * ncols + c] = acc[p][q];
    }
}

// -------------------------------------------------------------- K3: W finish ----
// W2[k][c] = sum_k' T[k'][k] * (sum_g Wp[g][k'][c]); 32 columns x 8 slices of 4 k per CTA.  The
// group sum is latency bound (a few loads per thread), so 8 groups x 4 k are loaded before any add.
constexpr int kWfCols = 32;
constexpr int kWqSlices = 8;
template <typename T>
__global__ void __launch_bounds__(256)
    qr_wfinish_kernel(const T* __restrict__ Wp, const T* __restrict__ Tm, T* __restrict__ W2, int ncols,
                      int ngroups) {
  __shared__ T Ts[kPB * kPB];
  __shared__ T wsum[kPB][kWfCols + 1];
  for (int e = threadIdx.x; e < kPB * kPB; e += blockDim.x) Ts[e] = Tm[e];
  const int cl = threadIdx.x & (kWfCols - 1), ks = threadIdx.x / kWfCols;  // ks: 0..7
  const int c = blockIdx.x * kWfCols + cl;
  if (c < ncols) {
    T w[4] = {T(0), T(0), T(0), T(0)};
    for (int g0 = 0; g0 < ngroups; g0 += 8) {
      T t[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          t[u][k] = g0 + u < ngroups ? Wp[((size_t)(g0 + u) * kPB + 4 * ks + k) * ncols + c] : T(0);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] += t[u][k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) wsum[4 * ks + k][cl] = w[k];
  }
  __syncthreads();
  if (c >= ncols) return;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int k = 4 * ks + kk;
    T s = T(0);
    for (int kp = 0; kp <= k; ++kp) s = fma_(Ts[kp * kPB + k], wsum[kp][cl], s);  // T upper triangular
    W2[(size_t)k * ncols + c] = s;
  }
}

// Pre-reduction of the row-group partials for the narrow inner updates of the two-level blocking: those use up
// to 592 row groups (one wave of the W kernel at 4 CTAs per SM) but only <= 96 columns, so qr_wfinish_kernel
// ran 3 CTAs that each walked all the groups serially (75 us per launch, latency bound).  Here
// `nslices` x (32 * ncols / 256) CTAs each sum a contiguous slice of the groups in a fixed order;
// qr_wfinish_kernel then sums the nslices results.  Wq[s][e] = sum_{g in slice s} Wp[g][e], e < 32 * ncols.
template <typename T>
__global__ void __launch_bounds__(256)
    qr_wpresum_kernel(const T* __restrict__ Wp, T* __restrict__ Wq, int elems, int ngroups, int per_slice) {
  const int e = blockIdx.x * 256 + threadIdx.x, sl = blockIdx.y;
  if (e >= elems) return;
  const int g_lo = sl * per_slice, g_hi = min(ngroups, g_lo + per_slice);
  T s = T(0);
  for (int g0 = g_lo; g0 < g_hi; g0 += 8) {
    T t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = g0 + u < g_hi ? Wp[(size_t)(g0 + u) * elems + e] : T(0);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += t[u];
  }
  Wq[(size_t)sl * elems + e] = s;
}

// ------------------------------------------------------------------ K4: update ----
// A[i][cbase + c] -= sum_k v(i,k) W2[k][c].
// Fallback for rows that are not 16-byte aligned: one 128 x 128 tile per CTA, thread tile 8 x 8.
constexpr int kUpRows = 128, kUpCols = 128, kUpLd = kUpRows + 4;
template <typename T>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : 1)
    qr_update_simple_kernel(T* __restrict__ a, const T* __restrict__ W2, int m, int n, int j0, int ncols) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Vt = reinterpret_cast<T*>(smem_raw);  // [32][132]: V tile transposed, padded against conflicts
  T* Ws = Vt + kPB * kUpLd;                // [32][128]
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, rblk = blockIdx.y;
  const int cbase = j0 + kPB + tile * kUpCols;
  const int rb = j0 + rblk * kUpRows;
  for (int idx = tid; idx < kUpRows * kPB; idx += 256) {
    const int r = idx / kPB, c = idx % kPB, gr = rb + r;
    Vt[c * kUpLd + r] = gr < m ? vmask<T>(a[(size_t)gr * n + j0 + c], gr, j0, c) : T(0);
  }
  for (int idx = tid; idx < kPB * kUpCols; idx += 256) {
    const int k = idx / kUpCols, c = idx % kUpCols;
    const int gc = tile * kUpCols + c;
    Ws[k * kUpCols + c] = gc < ncols ? W2[(size_t)k * ncols + gc] : T(0);
  }
  __syncthreads();
  const int ri = tid >> 4, ci = tid & 15;  // 16 row groups of 8, 16 column groups of 4 + 4
  T acc[8][8];
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[p][q] = T(0);
#pragma unroll 4
  for (int k = 0; k < kPB; ++k) {
    T v[8], w[8];
    lds4<T>(Vt + k * kUpLd + 8 * ri, v);
    lds4<T>(Vt + k * kUpLd + 8 * ri + 4, v + 4);
    lds4<T>(Ws + k * kUpCols + 4 * ci, w);
    lds4<T>(Ws + k * kUpCols + 64 + 4 * ci, w + 4);
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[p][q] = fma_(v[p], w[q], acc[p][q]);
  }
  using VT = typename V16K<T>::type;
  constexpr int V = 16 / sizeof(T);
  const bool vec = (n % V == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                   (tile * kUpCols + kUpCols <= ncols);
  if (vec) {
    // 4 rows at a time: all loads of the group are issued before the first dependent store
#pragma unroll
    for (int pg = 0; pg < 2; ++pg) {
      VT val[4][2][4 / V];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int gr = rb + 8 * ri + 4 * pg + p;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int u = 0; u < 4 / V; ++u)
            if (gr < m)
              val[p][h][u] = reinterpret_cast<const VT*>(a + (size_t)gr * n + cbase + 64 * h + 4 * ci)[u];
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int gr = rb + 8 * ri + 4 * pg + p;
        if (gr >= m) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int u = 0; u < 4 / V; ++u) {
            T* pv = reinterpret_cast<T*>(&val[p][h][u]);
#pragma unroll
            for (int e = 0; e < V; ++e) pv[e] = pv[e] - acc[4 * pg + p][4 * h + u * V + e];
            reinterpret_cast<VT*>(a + (size_t)gr * n + cbase + 64 * h + 4 * ci)[u] = val[p][h][u];
          }
      }
    }
  } else {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int gr = rb + 8 * ri + p;
      if (gr >= m) continue;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int cc = (q < 4 ? 4 * ci + q : 64 + 4 * ci + q - 4);
        const int gc = tile * kUpCols + cc;
        if (gc < ncols) {
          T* pp = a + (size_t)gr * n + cbase + cc;
          *pp = *pp - acc[p][q];
        }
      }
    }
  }
}

// Main path (16-byte aligned rows).  A CTA owns one 128-column tile and walks a strip of row
// blocks (RT = 128 rows fp32 / 64 rows fp64).  W2's tile stays in shared memory for the whole
// strip; per row block the A2 tile is fetched global -> shared by cp.async while the block's
// FMAs run (it is only needed by the epilogue), and the V tile of the NEXT block is fetched,
// transposed and XOR-swizzled, into the other half of a double buffer.  No global load is ever
// waited on in the FMA loop.  Thread tile MR x 8 (MR = 8 fp32 / 4 fp64).
template <typename T>
struct UpCfg {
  static constexpr int MR = sizeof(T) == 4 ? 8 : 4;
  static constexpr int RT = 16 * MR;
  static constexpr size_t smem = ((size_t)kPB * kUpCols + 2 * (size_t)kPB * RT + (size_t)RT * kUpCols) * sizeof(T);
};
__device__ __forceinline__ int up_swz(int k) { return (k & 7) << 2; }

template <typename T>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : 1)
    qr_update_kernel(T* __restrict__ a, const T* __restrict__ W2, int m, int n, int j0, int ncols,
                     int rblocks, int per_strip) {
  constexpr int MR = UpCfg<T>::MR, RT = UpCfg<T>::RT;
  constexpr int V = 16 / sizeof(T);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ws = reinterpret_cast<T*>(smem_raw);  // [32][128]
  T* Vt = Ws + kPB * kUpCols;              // 2 x [32][RT], element (k, r) at k*RT + (r ^ swz(k))
  T* At = Vt + 2 * kPB * RT;               // [RT][128]
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int cbase = j0 + kPB + tile * kUpCols;
  const int cw = min(kUpCols, ncols - tile * kUpCols);
  const int b0 = blockIdx.y * per_strip, b1 = min(rblocks, b0 + per_strip);
  if (b0 >= b1) return;
  for (int idx = tid; idx < kPB * kUpCols; idx += 256) {
    const int k = idx / kUpCols, c = idx % kUpCols;
    Ws[idx] = c < cw ? W2[(size_t)k * ncols + tile * kUpCols + c] : T(0);
  }
  auto issue_v = [&](int buf, int b) {
    T* dst = Vt + buf * (kPB * RT);
    const int rb = j0 + b * RT;
#pragma unroll
    for (int q = 0; q < kPB * RT / 256; ++q) {
      const int e = tid + q * 256;
      const int k = ((e >> 5) & 3) * 8 + (e & 7), r = (e >> 7) * 4 + ((e >> 3) & 3), gr = rb + r;
      const bool ok = gr < m;
      cp_async_zfill(dst + k * RT + (r ^ up_swz(k)), ok ? a + (size_t)gr * n + j0 + k : a, (int)sizeof(T),
                     ok ? (int)sizeof(T) : 0);
    }
  };
  auto issue_a = [&](int b) {
    const int rb = j0 + b * RT;
    constexpr int ppr = kUpCols / V;  // 16-byte pieces per row
#pragma unroll
    for (int q = 0; q < RT * ppr / 256; ++q) {
      const int pc = tid + q * 256, r = pc / ppr, c = (pc % ppr) * V, gr = rb + r;
      int nb = gr < m ? (cw - c) * (int)sizeof(T) : 0;
      nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
      cp_async_zfill(At + r * kUpCols + c, nb ? a + (size_t)gr * n + cbase + c : a, 16, nb);
    }
  };
  const int ri = tid >> 4, ci = tid & 15;  // 16 row groups of MR, 16 column groups of 4 + 4
  issue_v(0, b0);
  cp_async_commit();
  for (int b = b0; b < b1; ++b) {
    const int buf = (b - b0) & 1;
    __syncthreads();  // epilogue b-1 done with At, FMA loop b-1 done with the other V buffer
    issue_a(b);
    if (b + 1 < b1) issue_v(buf ^ 1, b + 1);
    cp_async_commit();
    cp_async_wait<1>();  // V(b) has landed
    __syncthreads();
    T* Vb = Vt + buf * (kPB * RT);
    if (b == 0) {  // rows crossing the panel's diagonal block: unit diagonal, zeros above
      for (int e = tid; e < kPB * kPB; e += 256) {
        const int k = e / kPB, r = e % kPB, gr = j0 + r;
        T* p = Vb + k * RT + (r ^ up_swz(k));
        *p = gr < m ? vmask<T>(*p, gr, j0, k) : T(0);
      }
      __syncthreads();
    }
    T acc[MR][8];
#pragma unroll
    for (int p = 0; p < MR; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[p][q] = T(0);
#pragma unroll 8
    for (int k = 0; k < kPB; ++k) {
      T v[MR], w[8];
#pragma unroll
      for (int h = 0; h < MR / 4; ++h) lds4<T>(Vb + k * RT + ((MR * ri + 4 * h) ^ up_swz(k)), v + 4 * h);
      lds4<T>(Ws + k * kUpCols + 4 * ci, w);
      lds4<T>(Ws + k * kUpCols + 64 + 4 * ci, w + 4);
#pragma unroll
      for (int p = 0; p < MR; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[p][q] = fma_(v[p], w[q], acc[p][q]);
    }
    cp_async_wait<0>();  // A2 tile of this block (and V of the next) have landed
    __syncthreads();
    const int rb = j0 + b * RT;
#pragma unroll
    for (int p = 0; p < MR; ++p) {
      const int gr = rb + MR * ri + p;
      if (gr >= m) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = 64 * h + 4 * ci;
        if (cc >= cw) continue;
        T t[4];
        lds4<T>(At + (MR * ri + p) * kUpCols + cc, t);
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] = t[e] - acc[p][4 * h + e];
        using VT = typename V16K<T>::type;
#pragma unroll
        for (int u = 0; u < 4 / V; ++u)  // cw is a multiple of V here: 16-byte pieces are all-or-nothing
          if (cc + u * V < cw)
            reinterpret_cast<VT*>(a + (size_t)gr * n + cbase + cc)[u] = reinterpret_cast<const VT*>(t)[u];
      }
    }
  }
}

// Tensor-core variant of the update (fp32, aligned rows; the default there, LXB_QR_TC=0 turns it off):
// P = V W on the 5th-generation tensor cores with the 3xTF32 split (x = hi + lo, hi = the 19 leading
// bits; P = Vhi Whi + Vhi Wlo + Vlo Whi accumulated in fp32 in tensor memory), so the products keep
// ~2^-21 relative accuracy instead of TF32's 2^-11.  One CTA of 4 warps owns a 128-column tile and
// walks a strip of 128-row blocks:
//   W tile  -> Bhi / Blo  (128 x 32, K-major, 128-byte swizzle; written once per CTA)
//   V block -> Ahi / Alo  (128 x 32, same layout; thread t stages row t)
//   one thread issues 12 tcgen05.mma (M = N = 128, K = 8) into 128 TMEM columns and commits to an
//   mbarrier; every warp then reads its 32 TMEM lanes back (tcgen05.ld 32x32b.x32) and does the
//   read-modify-write of A2, one 128-byte row segment per thread per 32-column chunk.
namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, SWIZZLE_128B operand tile with 128-byte rows: 8-row groups are 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
// byte offset of the 16-byte chunk `c` (0..7) of row `r` inside a swizzled 128 x 32 fp32 tile
__device__ __forceinline__ int sw128(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }
// x = hi + lo with BOTH parts rounded to nearest TF32 (cvt.rna): truncating them instead (the tensor core
// ignores the 13 low mantissa bits of an fp32 operand) biases every product towards zero by ~2^-22, and
// over a 262144-row contraction that bias does not average out (measured: ||R - R64||_F / ||R64||_F grew
// linearly with n, 2.5e-6 at n = 2048, against 1e-7 with rounding).
// cvt.rna.tf32.f32 is emulated on sm_100a (FSETP + VIADD + LOP3 + SEL per value: it was most of the staging
// instructions); round-to-nearest, ties away, on the magnitude bits is an add and a mask.
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_store(unsigned char* hi, unsigned char* lo, int r, int c, float4 x) {
  float4 h, l;
  h.x = tf32_rn(x.x); l.x = tf32_rn(x.x - h.x);
  h.y = tf32_rn(x.y); l.y = tf32_rn(x.y - h.y);
  h.z = tf32_rn(x.z); l.z = tf32_rn(x.z - h.z);
  h.w = tf32_rn(x.w); l.w = tf32_rn(x.w - h.w);
  *reinterpret_cast<float4*>(hi + sw128(r, c)) = h;
  *reinterpret_cast<float4*>(lo + sw128(r, c)) = l;
}
}  // namespace tc

constexpr int kTcThreads = 128;
constexpr size_t kTcSmem = 4 * 16384 + 1024 /* alignment slack */ + 64;

__global__ void __launch_bounds__(kTcThreads)
    qr_update_tc_kernel(float* __restrict__ a, const float* __restrict__ W2, int m, int n, int j0, int ncols,
                        int rblocks, int per_strip) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* Bhi = base;
  unsigned char* Blo = base + 16384;
  unsigned char* Ahi = base + 2 * 16384;
  unsigned char* Alo = base + 3 * 16384;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(base + 4 * 16384);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + 4 * 16384 + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int cbase = j0 + kPB + tile * 128;
  const int cw = min(128, ncols - tile * 128);
  const int b0 = blockIdx.y * per_strip, b1 = min(rblocks, b0 + per_strip);
  if (b0 >= b1) return;  // uniform per CTA, before any allocation
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tslot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // W tile: thread t owns column t (row t of the K-major B operand)
  {
    float w[kPB];
#pragma unroll
    for (int k = 0; k < kPB; ++k) w[k] = tid < cw ? W2[(size_t)k * ncols + tile * 128 + tid] : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) tc::split_store(Bhi, Blo, tid, c, make_float4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  const uint32_t a_hi = tc::smem_u32(Ahi), a_lo = tc::smem_u32(Alo), b_hi = tc::smem_u32(Bhi), b_lo = tc::smem_u32(Blo);
  uint32_t phase = 0;
  // The CTA is latency bound on its A2 read-modify-write (ncu: 12 warps per SM, long_scoreboard 18
  // per issue), so every block's A2 tile and V rows are pulled into L2 one block ahead.
  auto prefetch_block = [&](int bb) {
    const int prb = j0 + bb * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int line = tid + i * 128, r = line >> 2, seg = line & 3, gr = prb + r;  // 4 x 128-byte lines per row
      if (gr < m && seg * 32 < cw) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)gr * n + cbase + seg * 32));
    }
    if (prb + tid < m) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)(prb + tid) * n + j0));
  };
  prefetch_block(b0);
  for (int b = b0; b < b1; ++b) {
    const int rb = j0 + b * 128;
    if (b + 1 < b1) prefetch_block(b + 1);
    {
      // 8 lanes per row (one 16-byte chunk each), 16 rows per sweep: fully coalesced 128-byte rows
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (tid >> 3) + 16 * i, c = tid & 7, gr = rb + r;
        v[i] = gr < m ? reinterpret_cast<const float4*>(a + (size_t)gr * n + j0)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (tid >> 3) + 16 * i, c = tid & 7, gr = rb + r;
        if (b == 0) {  // rows crossing the panel's diagonal block: unit diagonal, zeros above
          v[i].x = vmask<float>(v[i].x, gr, j0, 4 * c);
          v[i].y = vmask<float>(v[i].y, gr, j0, 4 * c + 1);
          v[i].z = vmask<float>(v[i].z, gr, j0, 4 * c + 2);
          v[i].w = vmask<float>(v[i].w, gr, j0, 4 * c + 3);
        }
        tc::split_store(Ahi, Alo, r, c, v[i]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // K = 32 in steps of 8 tf32 = 32 bytes along the swizzled row
        const uint64_t dah = tc::umma_desc(a_hi + 32 * ks), dal = tc::umma_desc(a_lo + 32 * ks);
        const uint64_t dbh = tc::umma_desc(b_hi + 32 * ks), dbl = tc::umma_desc(b_lo + 32 * ks);
        tc::mma_tf32(tmem, dal, dbh, ks > 0 ? 1u : 0u);  // small terms first
        tc::mma_tf32(tmem, dah, dbl, 1u);
        tc::mma_tf32(tmem, dah, dbh, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
    }
    {
      uint32_t done = 0;
      for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(tc::smem_u32(mbar)), "r"(phase) : "memory");
      if (!done) __trap();  // bounded wait: fail LOUDLY (launch error) instead of updating with stale TMEM
      phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // Epilogue.  tcgen05.ld hands every thread 32 consecutive columns of ITS row; going to global
    // memory like that would touch 32 different lines per instruction, so each warp turns its
    // 32 x 32 chunk around in shared memory (the A tiles are dead once the MMAs have completed;
    // 16-byte chunks XOR-swizzled by row) and does the read-modify-write with 8 lanes per row.
    float4* S = reinterpret_cast<float4*>(Ahi + warp * 4096);  // 32 rows x 8 chunks
#pragma unroll 1
    for (int qc = 0; qc < 4; ++qc) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + 32 * qc;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __syncwarp();  // previous chunk's readers are done with S
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4)
        S[lane * 8 + (g4 ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1]),
                                                      __uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3]));
      __syncwarp();
      const int c = lane & 7, col = 32 * qc + 4 * c;
      float4 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3), gr = rb + 32 * warp + rr;
        if (gr < m && col < cw) x[i] = *reinterpret_cast<const float4*>(a + (size_t)gr * n + cbase + col);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3), gr = rb + 32 * warp + rr;
        if (gr < m && col < cw) {  // aligned rows: ncols (hence cw) is a multiple of 4
          const float4 p = S[rr * 8 + (c ^ (rr & 7))];
          x[i].x -= p.x; x[i].y -= p.y; x[i].z -= p.z; x[i].w -= p.w;
          *reinterpret_cast<float4*>(a + (size_t)gr * n + cbase + col) = x[i];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();  // TMEM and the A tiles are free for the next block
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// Transposed staging for the W = V^T A2 product on tcgen05 (qr_wbig_tc_kernel below): the contraction runs
// over the ROWS, so the 32-row chunks are TRANSPOSED while they are staged and both operands become K-major,
// the configuration of the update kernels.  (An MN-major encoding of the untransposed chunks in the plain
// SWIZZLE_128B layout returns zeros; tf32 needs SWIZZLE_128B_BASE32B -- tools/mn_major_probe.cu -- which
// is what qr_wbig_ws_kernel, the default for the trailing W, uses.  This kernel remains for the Gram
// matrix of V and as the LXB_QR_W3=0 fallback.)  Thread (rq = tid & 7, cq) loads a 4 x 4 block (rows 4rq.., columns 4cq..),
// transposes it in registers and stores four 16-byte K-chunks; a quarter warp covers the 8 chunks of one
// tile row, so the stores are conflict free.  (A K = 32 version of this kernel, one panel at a time, was
// validated in situ this round and measured no faster than the SIMT W kernel -- 317 vs 313 ms -- and removed.)
namespace tc {
__device__ __forceinline__ void split_store_t(unsigned char* hi, unsigned char* lo, int col0, int rq, const float4 (&x)[4]) {
  split_store(hi, lo, col0 + 0, rq, make_float4(x[0].x, x[1].x, x[2].x, x[3].x));
  split_store(hi, lo, col0 + 1, rq, make_float4(x[0].y, x[1].y, x[2].y, x[3].y));
  split_store(hi, lo, col0 + 2, rq, make_float4(x[0].z, x[1].z, x[2].z, x[3].z));
  split_store(hi, lo, col0 + 3, rq, make_float4(x[0].w, x[1].w, x[2].w, x[3].w));
}
}  // namespace tc

constexpr int kWtcRows = 32;  // rows per chunk

// ------------------------------------------------------------ two-level blocking ----
// With 32-column panels every trailing pass moves the whole trailing matrix through HBM three times for
// 64 flops per element: 128 passes x 3 x ~2.1 GB = 0.8 TB per factorisation, a ~125 ms floor.
// The fp32 path therefore groups FOUR panels into a 128-column outer block: inside the block the 32-wide
// updates above touch at most 96 columns, and the trailing matrix is updated ONCE per outer block with
// the 128 reflectors of the block, A2 <- A2 - V (T^T (V^T A2)), T = larft(V, taus) (128 x 128):
//   qr_wbig_tc_kernel     W = V^T A2 on tcgen05 (3xTF32), M = 128 columns of A2, N = 128 reflectors, the
//                         32-row chunks transposed while staged (both operands K-major), partial sums per
//                         row group accumulated in TMEM; with MASKA the "A2" tile is the V block itself and
//                         the kernel returns the Gram matrix G = V^T V the T factor needs
//   qr_tbig_kernel        G = sum of the partials; T by LAPACK's larft recurrence T(0:i,i) = -tau_i T G(0:i,i)
//   qr_wfinish128_kernel  W = sum of the row-group partials; Y = T^T W, stored TRANSPOSED (column-major in K)
//                         so that it is a K-major B operand
//   qr_update128_tc_kernel A2 -= V Y on tcgen05: 128-row x 64-column tiles, K = 128 as four staged 32-wide
//                         chunks accumulated in one TMEM tile, the Y tile resident in shared memory
namespace tc {
constexpr uint32_t kIdesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // N = 64
__device__ __forceinline__ void mma_tf32_n64(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kIdesc64), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* mbar, uint32_t& phase) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 24) && !done; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
  if (!done) __trap();  // bounded wait: fail loudly, never continue with stale TMEM
  phase ^= 1;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace tc

constexpr int kOB = 128;  // outer block width (reflectors per trailing update)
constexpr size_t kWbigSmem = 4 * 16384 + 1024 + 64;

// Wp[grp][k][c] = sum over the rows of group grp of V[r][k] * A2[r][c], k < 128, c < ncols;
// A2 = columns cbase0 .. cbase0 + ncols - 1 of `a` (MASKA: cbase0 == j0, the V block itself, masked).
template <bool MASKA>
__global__ void __launch_bounds__(kTcThreads, 2)
    qr_wbig_tc_kernel(const float* __restrict__ a, float* __restrict__ Wp, int m, int n, int j0, int cbase0,
                      int ncols, int ngroups) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* Ahi = base;                  // 128 tile rows (A2 columns) x 128 B (32 rows of the chunk)
  unsigned char* Alo = base + 16384;
  unsigned char* Bhi = base + 2 * 16384;      // 128 tile rows (reflectors) x 128 B
  unsigned char* Blo = base + 3 * 16384;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(base + 4 * 16384);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + 4 * 16384 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = blockIdx.x, grp = blockIdx.y;
  const int cbase = cbase0 + tile * 128;
  const int cw = min(128, ncols - tile * 128);
  const int rows_total = m - j0;
  const int per = (((rows_total + ngroups - 1) / ngroups) + kWtcRows - 1) / kWtcRows * kWtcRows;
  const int r0 = j0 + grp * per, r1 = min(m, r0 + per);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tslot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  const uint32_t a_hi = tc::smem_u32(Ahi), a_lo = tc::smem_u32(Alo), b_hi = tc::smem_u32(Bhi), b_lo = tc::smem_u32(Blo);
  // staging map: a 32-row chunk of 128 columns = 8 row quads x 32 column quads of 4 x 4 blocks, two per thread
  const int rq = tid & 7;
  float4 pa[2][4], pv[2][4];
  auto fetch = [&](int rb) {
    const bool diag = rb < j0 + kOB;  // rows crossing the block's diagonal: unit diagonal, zeros above
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int cq = (tid >> 3) + 16 * i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gr = rb + 4 * rq + j;
        float4 x = (gr < r1 && 4 * cq < cw) ? reinterpret_cast<const float4*>(a + (size_t)gr * n + cbase)[cq]
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v = gr < r1 ? reinterpret_cast<const float4*>(a + (size_t)gr * n + j0)[cq] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (diag && gr < r1) {
          v.x = vmask<float>(v.x, gr, j0, 4 * cq);
          v.y = vmask<float>(v.y, gr, j0, 4 * cq + 1);
          v.z = vmask<float>(v.z, gr, j0, 4 * cq + 2);
          v.w = vmask<float>(v.w, gr, j0, 4 * cq + 3);
          if (MASKA) {
            const int c0 = cbase - j0 + 4 * cq;
            x.x = vmask<float>(x.x, gr, j0, c0);
            x.y = vmask<float>(x.y, gr, j0, c0 + 1);
            x.z = vmask<float>(x.z, gr, j0, c0 + 2);
            x.w = vmask<float>(x.w, gr, j0, c0 + 3);
          }
        }
        pa[i][j] = x;
        pv[i][j] = v;
      }
    }
  };
  const int nchunks = r1 > r0 ? (r1 - r0 + kWtcRows - 1) / kWtcRows : 0;
  uint32_t phase = 0;
  // The tensor core accumulates into TMEM with truncation, so a long running sum there picks up a
  // systematic bias (measured: ||R - R64||_F / ||R64||_F = 2.5e-6 at n = 2048, growing linearly with n, when
  // a whole row group -- ~280 chunks -- was summed in TMEM; 7e-8 with short sums).  At most kDrainEvery
  // 32-row chunks (48 MMAs) are accumulated in TMEM; the running sum over the row group lives in fp32
  // registers (round-to-nearest adds), 128 per thread.
  constexpr int kDrainEvery = 4;
  float acc[kOB];
#pragma unroll
  for (int k = 0; k < kOB; ++k) acc[k] = 0.f;
  auto drain = [&]() {
    tc::mbar_wait_or_trap(mbar, phase);
    asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
    for (int qc = 0; qc < 4; ++qc) {
      uint32_t r[32];
      tc::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 32 * qc, r);
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[32 * qc + k] += __uint_as_float(r[k]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  };
  // ncu: 69 % of the stall samples are long-scoreboard waits on the register fetch of the next chunk (it
  // is only one chunk ahead and HBM latency under load exceeds a chunk's staging + MMA time), so the lines
  // of the chunk kPfAhead further on are pulled into L2 now: 128 lines of A2 and 128 of V, one each per thread
  constexpr int kPfAhead = 6;
  auto prefetch_chunk = [&](int rb) {
    const int gr = rb + (tid >> 2), seg = tid & 3;
    if (gr < r1) {
      if (seg * 32 < cw) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)gr * n + cbase + seg * 32));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)gr * n + j0 + seg * 32));
    }
  };
  for (int c = 1; c <= kPfAhead && c < nchunks; ++c) prefetch_chunk(r0 + c * kWtcRows);
  if (nchunks > 0) fetch(r0);
  for (int ch = 0; ch < nchunks; ++ch) {
    if (ch + 1 + kPfAhead < nchunks) prefetch_chunk(r0 + (ch + 1 + kPfAhead) * kWtcRows);
    const bool fresh = ch % kDrainEvery == 0;  // this chunk starts a new TMEM sum
    if (ch > 0) {
      if (fresh) drain();                              // wait + add the finished TMEM sum to the registers
      else tc::mbar_wait_or_trap(mbar, phase);         // wait only: the previous chunk's tiles are free
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      tc::split_store_t(Ahi, Alo, 4 * ((tid >> 3) + 16 * i), rq, pa[i]);
      tc::split_store_t(Bhi, Blo, 4 * ((tid >> 3) + 16 * i), rq, pv[i]);
    }
    if (ch + 1 < nchunks) fetch(r0 + (ch + 1) * kWtcRows);  // in flight while the MMAs run
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int ks = 0; ks < kWtcRows / 8; ++ks) {
        const uint64_t dah = tc::umma_desc(a_hi + 32 * ks), dal = tc::umma_desc(a_lo + 32 * ks);
        const uint64_t dbh = tc::umma_desc(b_hi + 32 * ks), dbl = tc::umma_desc(b_lo + 32 * ks);
        tc::mma_tf32(tmem, dal, dbh, (!fresh || ks > 0) ? 1u : 0u);  // small terms first
        tc::mma_tf32(tmem, dah, dbl, 1u);
        tc::mma_tf32(tmem, dah, dbh, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
    }
  }
  if (nchunks > 0) drain();
  // lane of TMEM = column of the A2 tile, TMEM column = reflector k
  const int col = tile * 128 + tid;
  if (tid < cw) {
    float* out = Wp + ((size_t)grp * kOB) * ncols + col;
#pragma unroll
    for (int k = 0; k < kOB; ++k) out[(size_t)k * ncols] = acc[k];
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// G = sum of the row-group partials of the Gram kernel (fixed order, 8 loads in flight per thread)
__global__ void __launch_bounds__(256) qr_gsum_kernel(const float* __restrict__ Gp, float* __restrict__ G, int ngroups) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  float s = 0.f;
  for (int g0 = 0; g0 < ngroups; g0 += 8) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = g0 + u < ngroups ? Gp[(size_t)(g0 + u) * kOB * kOB + e] : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += t[u];
  }
  G[e] = s;
}

// G = V^T V from the row-group partials (ncols = 128), then T (128 x 128, upper triangular) by LAPACK's larft:
// T(i,i) = tau_i, T(0:i, i) = -tau_i * T(0:i, 0:i) * G(0:i, i).  One CTA.
constexpr int kTbigThreads = 512;
__global__ void __launch_bounds__(kTbigThreads) qr_tbig_kernel(const float* __restrict__ Gp, const float* __restrict__ taus,
                                                              float* __restrict__ Tb, int ngroups) {
  extern __shared__ float tb_smem[];
  float* G = tb_smem;                 // [128][129]
  float* Ts = tb_smem + kOB * (kOB + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < kOB * kOB; e += kTbigThreads) {
    const int k = e / kOB, c = e % kOB;
    float s = 0.f;
    for (int g = 0; g < ngroups; ++g) s += Gp[((size_t)g * kOB + k) * kOB + c];
    G[k * (kOB + 1) + c] = s;
    Ts[k * (kOB + 1) + c] = 0.f;
  }
  __syncthreads();
  for (int i = 0; i < kOB; ++i) {
    const float tau = taus[i];
    float z = 0.f;
    if (tid < i) {
      for (int k = tid; k < i; ++k) z = fmaf(Ts[tid * (kOB + 1) + k], G[k * (kOB + 1) + i], z);
    }
    __syncthreads();
    if (tid < i) Ts[tid * (kOB + 1) + i] = -tau * z;
    if (tid == i) Ts[i * (kOB + 1) + i] = tau;
    __syncthreads();
  }
  for (int e = tid; e < kOB * kOB; e += kTbigThreads) Tb[e] = Ts[(e / kOB) * (kOB + 1) + e % kOB];
}

// W = sum of the row-group partials; Yt[c][i] = sum_k T[k][i] W[k][c]  (Y = T^T W, stored K-major per column)
constexpr int kWf128Cols = 32;
__global__ void __launch_bounds__(256) qr_wfinish128_kernel(const float* __restrict__ Wp, const float* __restrict__ Tb,
                                                            float* __restrict__ Yt, int ncols, int ngroups) {
  __shared__ float Ws[kOB][kWf128Cols + 1];
  const int tid = threadIdx.x, c0 = blockIdx.x * kWf128Cols;
  for (int e = tid; e < kOB * kWf128Cols; e += 256) {
    const int k = e / kWf128Cols, c = e % kWf128Cols;
    float s = 0.f;
    if (c0 + c < ncols)
      for (int g = 0; g < ngroups; ++g) s += Wp[((size_t)g * kOB + k) * ncols + c0 + c];
    Ws[k][c] = s;
  }
  __syncthreads();
  for (int e = tid; e < kOB * kWf128Cols; e += 256) {
    const int i = e % kOB, c = e / kOB;
    if (c0 + c >= ncols) continue;
    float y = 0.f;
    for (int k = 0; k <= i; ++k) y = fmaf(Tb[k * kOB + i], Ws[k][c], y);
    Yt[(size_t)(c0 + c) * kOB + i] = y;
  }
}

// A2 -= V Y for the 128 reflectors of an outer block; Yt[c][k] (K-major per column).
constexpr size_t kUp128Smem = 8 * 8192 + 2 * 16384 + 1024 + 64;
__global__ void __launch_bounds__(kTcThreads, 2)
    qr_update128_tc_kernel(float* __restrict__ a, const float* __restrict__ Yt, int m, int n, int j0, int ncols,
                           int rblocks, int per_strip) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* Bt = base;                   // 4 K-chunks x {hi, lo} x (64 columns x 128 B)
  unsigned char* Ahi = base + 8 * 8192;       // 128 rows x 128 B (one K-chunk of V)
  unsigned char* Alo = Ahi + 16384;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(Alo + 16384);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(Alo + 16384 + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int cbase = j0 + kOB + tile * 64;
  const int cw = min(64, ncols - tile * 64);
  const int b0 = blockIdx.y * per_strip, b1 = min(rblocks, b0 + per_strip);
  if (b0 >= b1) return;  // uniform per CTA, before any allocation
  if (warp == 0) {
    // four accumulators of 64 columns, one per K-chunk: TMEM accumulation truncates, so no sum there is
    // longer than the 12 MMAs of one chunk; the four are added in registers (round to nearest)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tslot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // Y tile: tile row = column of A2 (N index), 128 B of K per chunk
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int idx = tid + 128 * i, q = idx & 7, c = (idx >> 3) & 63, kc = idx >> 9;
    const float4 y = c < cw ? reinterpret_cast<const float4*>(Yt + (size_t)(tile * 64 + c) * kOB + 32 * kc)[q]
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    tc::split_store(Bt + (2 * kc) * 8192, Bt + (2 * kc + 1) * 8192, c, q, y);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  const uint32_t a_hi = tc::smem_u32(Ahi), a_lo = tc::smem_u32(Alo), b_t = tc::smem_u32(Bt);
  uint32_t phase = 0;
  // The CTA is latency bound on its global loads (ncu: long_scoreboard 52 %, 8 warps per SM), so (i) the
  // next row block's V rows and A2 tile are pulled into L2 one block ahead and (ii) the V chunk of step
  // kc + 1 is loaded into registers while the MMAs of step kc run.
  auto prefetch_block = [&](int bb) {
    const int prb = j0 + bb * 128;
    if (prb + tid < m) {
      const float* vr = a + (size_t)(prb + tid) * n + j0;
#pragma unroll
      for (int l = 0; l < 4; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(vr + 32 * l));
      if (cw > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)(prb + tid) * n + cbase));
      if (cw > 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)(prb + tid) * n + cbase + 32));
    }
  };
  auto load_v = [&](int rb, int kc, float4 (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = (tid >> 3) + 16 * i, c = tid & 7, gr = rb + r;
      v[i] = gr < m ? reinterpret_cast<const float4*>(a + (size_t)gr * n + j0 + 32 * kc)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  prefetch_block(b0);
  float4 v[8];
  load_v(j0 + b0 * 128, 0, v);
  for (int b = b0; b < b1; ++b) {
    const int rb = j0 + b * 128;
    if (b + 1 < b1) prefetch_block(b + 1);
#pragma unroll 1
    for (int kc = 0; kc < 4; ++kc) {
      if (kc > 0) tc::mbar_wait_or_trap(mbar, phase);  // the previous chunk's MMAs have read the A tiles
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (tid >> 3) + 16 * i, c = tid & 7, gr = rb + r;
        if (b == 0) {  // rows crossing the block's diagonal: unit diagonal, zeros above
          v[i].x = vmask<float>(v[i].x, gr, j0, 32 * kc + 4 * c);
          v[i].y = vmask<float>(v[i].y, gr, j0, 32 * kc + 4 * c + 1);
          v[i].z = vmask<float>(v[i].z, gr, j0, 32 * kc + 4 * c + 2);
          v[i].w = vmask<float>(v[i].w, gr, j0, 32 * kc + 4 * c + 3);
        }
        tc::split_store(Ahi, Alo, r, c, v[i]);
      }
      // next chunk (or the first chunk of the next row block) in flight while the tensor core works
      if (kc < 3) load_v(rb, kc + 1, v);
      else if (b + 1 < b1) load_v(rb + 128, 0, v);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t b_hi = b_t + (2 * kc) * 8192, b_lo = b_t + (2 * kc + 1) * 8192;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t dah = tc::umma_desc(a_hi + 32 * ks), dal = tc::umma_desc(a_lo + 32 * ks);
          const uint64_t dbh = tc::umma_desc(b_hi + 32 * ks), dbl = tc::umma_desc(b_lo + 32 * ks);
          tc::mma_tf32_n64(tmem + 64 * kc, dal, dbh, ks > 0 ? 1u : 0u);
          tc::mma_tf32_n64(tmem + 64 * kc, dah, dbl, 1u);
          tc::mma_tf32_n64(tmem + 64 * kc, dah, dbh, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
      }
    }
    tc::mbar_wait_or_trap(mbar, phase);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // epilogue: as in qr_update_tc_kernel, 32 x 32 chunks turned around in (dead) A tiles
    float4* S = reinterpret_cast<float4*>(Ahi + warp * 4096);
#pragma unroll 1
    for (int qc = 0; qc < 2; ++qc) {
      float p32[32];
      {
        uint32_t r[32];
        tc::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 32 * qc, r);
#pragma unroll
        for (int k = 0; k < 32; ++k) p32[k] = __uint_as_float(r[k]);
#pragma unroll
        for (int kc = 1; kc < 4; ++kc) {
          tc::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 64 * kc + 32 * qc, r);
#pragma unroll
          for (int k = 0; k < 32; ++k) p32[k] += __uint_as_float(r[k]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4)
        S[lane * 8 + (g4 ^ (lane & 7))] = make_float4(p32[4 * g4], p32[4 * g4 + 1], p32[4 * g4 + 2], p32[4 * g4 + 3]);
      __syncwarp();
      const int c = lane & 7, col = 32 * qc + 4 * c;
      float4 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3), gr = rb + 32 * warp + rr;
        if (gr < m && col < cw) x[i] = *reinterpret_cast<const float4*>(a + (size_t)gr * n + cbase + col);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3), gr = rb + 32 * warp + rr;
        if (gr < m && col < cw) {
          const float4 p = S[rr * 8 + (c ^ (rr & 7))];
          x[i].x -= p.x; x[i].y -= p.y; x[i].z -= p.z; x[i].w -= p.w;
          *reinterpret_cast<float4*>(a + (size_t)gr * n + cbase + col) = x[i];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();  // TMEM and the A tiles are free for the next block
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// Warp-specialised version of the K = 128 update (default; LXB_QR_UP2=0 selects qr_update128_tc_kernel).
// ncu showed the kernel above latency bound (long_scoreboard 52 %, tensor pipe 16 %): every thread stages,
// waits for the MMAs and writes back in turn, and the hi/lo split of the SAME V block is redone by every
// 64-column tile (12 of every 14 executed instructions).  So
//   qr_vsplit_kernel  splits V ONCE per outer block into the exact shared-memory operand images
//                     (swizzled 128 x 32 tiles, hi then lo: 32 KB per K-chunk) in the workspace, and
//   qr_update128_ws_kernel runs three jobs concurrently in one persistent CTA per SM:
//     warp 9     PRODUCER: one thread copies the 16 KB hi / lo images global -> shared with cp.async.bulk
//                (mbarrier complete_tx) into a 4-slot ring; the other lanes prefetch A2 tiles into L2
//     warp 8     one elected thread issues the tcgen05.mma (M128 N128 K8) of an image as soon as its slot is
//                full, commits the slot back to the producer and, after the 8th image, the accumulator
//     warps 0-7  EPILOGUE: A2 tile (128 x 128) loaded BEFORE the accumulator is waited for, TMEM -> registers,
//                turned around in shared memory, subtracted and stored; two TMEM accumulators (2 x 128
//                columns) let the MMAs of row block b + 1 run under the epilogue of block b.
// All waits are mbarrier waits with a bounded spin that traps (no hang on a protocol error).
constexpr int kU2Threads = 320;  // warps 0-7 epilogue, 8 MMA issuer, 9 producer
constexpr int kU2Slots = 4;      // ring of 16 KB operand images
constexpr size_t kU2Smem = 8 * 16384 + kU2Slots * 16384 + 8 * 4096 + 1024 + 128;
constexpr size_t kVimgBlockFloats = 4 * 2 * 4096;  // one 128-row block: 4 K-chunks x {hi, lo} x 16 KB

namespace tc {
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  if (!done) __trap();
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldg128(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stg128(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// split_store with shared-window addresses (STS, not generic ST)
__device__ __forceinline__ void split_sts(uint32_t hi, uint32_t lo, int r, int c, float4 x) {
  float4 h, l;
  h.x = tf32_rn(x.x); l.x = tf32_rn(x.x - h.x);
  h.y = tf32_rn(x.y); l.y = tf32_rn(x.y - h.y);
  h.z = tf32_rn(x.z); l.z = tf32_rn(x.z - h.z);
  h.w = tf32_rn(x.w); l.w = tf32_rn(x.w - h.w);
  sts128(hi + sw128(r, c), h);
  sts128(lo + sw128(r, c), l);
}
}  // namespace tc

// V (columns j0 .. j0+127, rows j0 .., unit diagonal / zeros above applied) -> operand images:
//   vimg   K-major SWIZZLE_128B tiles [128 rows][32 k] (A operand of the update, M = rows)
//   vimgw  MN-major SWIZZLE_128B_BASE32B slabs [32 rows][128 k] (B operand of W = V^T A2, N = k, K = rows):
//          the only shared-memory layout tcgen05 accepts for a transposed tf32 operand -- column blocks of
//          32 k (4 KB apart = LBO), groups of 4 rows (512 B apart = SBO), 128 B rows whose 32-byte chunks are
//          XORed with (row & 3).  tools/mn_major_probe.cu checks the layout against a host product.
// grid (4 K-chunks, row blocks), 256 threads: thread = (row r0 + 32 i, 16-byte chunk q).
__device__ __forceinline__ int mn32_off(int rs, int cb, int q) {  // row rs of a 32-row slab, column block cb, 16 B chunk q
  return cb * 4096 + (rs >> 2) * 512 + (rs & 3) * 128 + ((((q >> 1) ^ (rs & 3)) << 5) | ((q & 1) << 4));
}
__global__ void __launch_bounds__(256) qr_vsplit_kernel(const float* __restrict__ a, float* __restrict__ vimg,
                                                        float* __restrict__ vimgw, int m, int n, int j0) {
  const int kc = blockIdx.x, b = blockIdx.y, q = threadIdx.x & 7, r0 = threadIdx.x >> 3;
  unsigned char* hi = reinterpret_cast<unsigned char*>(vimg + (size_t)b * kVimgBlockFloats + (size_t)kc * 8192);
  unsigned char* lo = hi + 16384;
  unsigned char* wimg = reinterpret_cast<unsigned char*>(vimgw + (size_t)b * kVimgBlockFloats);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + 32 * i, gr = j0 + b * 128 + r;
    float4 v = gr < m ? *reinterpret_cast<const float4*>(a + (size_t)gr * n + j0 + 32 * kc + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (b == 0) {
      v.x = vmask<float>(v.x, gr, j0, 32 * kc + 4 * q);
      v.y = vmask<float>(v.y, gr, j0, 32 * kc + 4 * q + 1);
      v.z = vmask<float>(v.z, gr, j0, 32 * kc + 4 * q + 2);
      v.w = vmask<float>(v.w, gr, j0, 32 * kc + 4 * q + 3);
    }
    float4 h, l;
    h.x = tc::tf32_rn(v.x); l.x = tc::tf32_rn(v.x - h.x);
    h.y = tc::tf32_rn(v.y); l.y = tc::tf32_rn(v.y - h.y);
    h.z = tc::tf32_rn(v.z); l.z = tc::tf32_rn(v.z - h.z);
    h.w = tc::tf32_rn(v.w); l.w = tc::tf32_rn(v.w - h.w);
    *reinterpret_cast<float4*>(hi + tc::sw128(r, q)) = h;
    *reinterpret_cast<float4*>(lo + tc::sw128(r, q)) = l;
    unsigned char* slab = wimg + (r >> 5) * 32768;  // slab = 32 rows: hi 16 KB, lo 16 KB
    *reinterpret_cast<float4*>(slab + mn32_off(r & 31, kc, q)) = h;
    *reinterpret_cast<float4*>(slab + 16384 + mn32_off(r & 31, kc, q)) = l;
  }
}

// W = V^T A2 (row-group partials), warp specialised like the update below and WITHOUT transposed staging:
// both operands are fed MN-major (rows = K), so a slab of 32 rows x 128 columns of A2 is split hi/lo
// where it lands (the transposing register shuffles of qr_wbig_tc_kernel were most of its instructions), and
// the V operand is a bulk copy of the pre-split slab image.
//   warps 0-3   DRAIN: every 4 slabs (48 MMAs -- longer TMEM sums pick up the accumulator's truncation bias)
//               the finished accumulator is added to a 128 x 128 fp32 tile in shared memory
//   warps 4-11  PRODUCERS: A2 slab global -> registers (two slabs in flight) -> hi/lo -> stage
//   warp 12     one elected thread issues 12 tcgen05.mma (M128 N128 K8) per slab into one of two accumulators
//   warp 13     bulk copies of the V slab images; other lanes prefetch A2 rows into L2
// Wp[grp][k][c], grp = blockIdx.y = a range of `pb` row blocks; lane of TMEM = column c, TMEM column = k.
constexpr int kW3Threads = 448;
constexpr size_t kW3Smem = 65536 + 2 * 65536 + 1024 + 128;
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr) {  // MN-major, SWIZZLE_128B_BASE32B, LBO 4096, SBO 512
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)1 << 61);
}
constexpr uint32_t kIdescMN = tc::kIdesc | (1u << 15) | (1u << 16);  // both operands transposed (MN-major)
__device__ __forceinline__ void mma_tf32_mn(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kIdescMN), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kW3Threads, 1)
    qr_wbig_ws_kernel(const float* __restrict__ a, const float* __restrict__ vimgw, float* __restrict__ Wp, int m, int n, int j0,
                      int cbase0, int ncols, int rblocks, int pb) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t accs = sbase;            // fp32 [128 k][128 c]
  const uint32_t stg = sbase + 65536;     // 2 stages x {A hi, A lo, B hi, B lo} x 16 KB
  unsigned char* gbase = smem_dyn + (sbase - tc::smem_u32(smem_dyn));
  uint64_t* fullA = reinterpret_cast<uint64_t*>(gbase + 65536 + 2 * 65536);
  uint64_t* fullB = fullA + 2;
  uint64_t* empty = fullB + 2;
  uint64_t* accf = empty + 2;
  uint64_t* acce = accf + 2;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(acce + 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, grp = blockIdx.y;
  const int cbase = cbase0 + tile * 128;
  const int cw = min(128, ncols - tile * 128);
  const int b0 = grp * pb, b1 = min(rblocks, b0 + pb);
  const int nslabs = (b1 - b0) * 4;  // > 0: the host launches ceil(rblocks / pb) groups
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tslot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int q = 0; q < 2; ++q) {
      tc::mbar_init(fullA + q, 256);
      tc::mbar_init(fullB + q, 1);
      tc::mbar_init(empty + q, 1);
      tc::mbar_init(accf + q, 1);
      tc::mbar_init(acce + q, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = tid; i < 4096; i += kW3Threads) tc::sts128(accs + 16 * i, make_float4(0.f, 0.f, 0.f, 0.f));
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  // slab s: stage s & 1 (use s >> 1); accumulator group g = s >> 2 in TMEM buffer g & 1 (use g >> 1)

  if (warp >= 4 && warp < 12) {
    // ---------------- producers ----------------
    const int pt = tid - 128;            // 0..255
    const int c4 = pt & 31, rr = pt >> 5;  // float4 column c4 of rows rr + 8 i
    const int cb = c4 >> 3, q = c4 & 7;
    const bool cok = 4 * c4 < cw;
    auto load_a = [&](int s, float4 (&v)[4]) {
      const int rb = j0 + b0 * 128 + s * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gr = rb + rr + 8 * i;
        v[i] = (s < nslabs && cok && gr < m) ? tc::ldg128(a + (size_t)gr * n + cbase + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto push = [&](int s, const float4 (&v)[4]) {
      const int st = s & 1, use = s >> 1;
      if (use > 0) tc::mbar_wait_parity(empty + st, (uint32_t)((use - 1) & 1));
      const uint32_t Ahi = stg + st * 65536, Alo = Ahi + 16384;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int off = mn32_off(rr + 8 * i, cb, q);
        float4 h, l;
        h.x = tc::tf32_rn(v[i].x); l.x = tc::tf32_rn(v[i].x - h.x);
        h.y = tc::tf32_rn(v[i].y); l.y = tc::tf32_rn(v[i].y - h.y);
        h.z = tc::tf32_rn(v[i].z); l.z = tc::tf32_rn(v[i].z - h.z);
        h.w = tc::tf32_rn(v[i].w); l.w = tc::tf32_rn(v[i].w - h.w);
        tc::sts128(Ahi + off, h);
        tc::sts128(Alo + off, l);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc::mbar_arrive(fullA + st);
    };
    float4 va[4], vb[4];
    load_a(0, va);
    load_a(1, vb);
    for (int s = 0; s < nslabs; s += 2) {  // nslabs is a multiple of 4
      push(s, va);
      load_a(s + 2, va);
      push(s + 1, vb);
      load_a(s + 3, vb);
    }
  } else if (warp == 13) {
    // ---------------- V slab images (bulk copies) + L2 prefetch of A2 ----------------
    auto prefetch_block = [&](int bb) {
      if (bb < b1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int gr = j0 + bb * 128 + lane + 32 * i;
          if (gr < m) {
            const float* row = a + (size_t)gr * n + cbase;
#pragma unroll
            for (int l = 0; l < 4; ++l)
              if (cw > 32 * l) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 32 * l));
          }
        }
      }
    };
    prefetch_block(b0 + 1);
    prefetch_block(b0 + 2);
    for (int s = 0; s < nslabs; ++s) {
      if (lane == 0) {
        const int st = s & 1, use = s >> 1;
        if (use > 0) tc::mbar_wait_parity(empty + st, (uint32_t)((use - 1) & 1));
        const float* src = vimgw + (size_t)(b0 + (s >> 2)) * kVimgBlockFloats + (size_t)(s & 3) * 8192;
        const uint32_t mb = tc::smem_u32(fullB + st);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(32768) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         stg + st * 65536 + 32768),
                     "l"(src), "r"(32768), "r"(mb)
                     : "memory");
      }
      if ((s & 3) == 0) prefetch_block(b0 + (s >> 2) + 3);
      __syncwarp();
    }
  } else if (warp == 12) {
    // ---------------- MMA issuer (whole warp, one elected lane issues) ----------------
    for (int s = 0; s < nslabs; ++s) {
      const int st = s & 1, use = s >> 1, g = s >> 2, buf = g & 1, ub = g >> 1;
      if ((s & 3) == 0 && ub > 0) tc::mbar_wait_parity(acce + buf, (uint32_t)((ub - 1) & 1));  // accumulator drained
      tc::mbar_wait_parity(fullA + st, (uint32_t)(use & 1));
      tc::mbar_wait_parity(fullB + st, (uint32_t)(use & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      if (tc::elect_one()) {
        const uint32_t sa = stg + st * 65536;
        const uint64_t dah = umma_desc_mn32(sa), dal = umma_desc_mn32(sa + 16384);
        const uint64_t dbh = umma_desc_mn32(sa + 32768), dbl = umma_desc_mn32(sa + 49152);
        const uint32_t acc = tmem + 128 * buf;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // 8 rows of K = 1 KB further on: + 64 in the address field
          mma_tf32_mn(acc, dal + 64 * ks, dbh + 64 * ks, ((s & 3) != 0 || ks > 0) ? 1u : 0u);  // small terms first
          mma_tf32_mn(acc, dah + 64 * ks, dbl + 64 * ks, 1u);
          mma_tf32_mn(acc, dah + 64 * ks, dbh + 64 * ks, 1u);
        }
        tc::umma_commit(empty + st);
        if ((s & 3) == 3) tc::umma_commit(accf + buf);
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    // ---------------- drain: TMEM accumulator -> += shared-memory tile ----------------
    const int ngrp = nslabs >> 2;
    for (int g = 0; g < ngrp; ++g) {
      const int buf = g & 1, ub = g >> 1;
      tc::mbar_wait_parity(accf + buf, (uint32_t)(ub & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int qc = 0; qc < 4; ++qc) {
        uint32_t r[32];
        tc::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 128 * buf + 32 * qc, r);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const uint32_t ad = accs + 4 * ((32 * qc + k) * 128 + tid);
          float t;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(ad) : "memory");
          t += __uint_as_float(r[k]);
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(ad), "f"(t) : "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      tc::mbar_arrive(acce + buf);
    }
    if (tid < cw) {
      float* out = Wp + ((size_t)grp * kOB) * ncols + tile * 128 + tid;
#pragma unroll 8
      for (int k = 0; k < kOB; ++k) {
        float t;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(accs + 4 * (k * 128 + tid)) : "memory");
        out[(size_t)k * ncols] = t;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

__global__ void __launch_bounds__(kU2Threads, 1)
    qr_update128_ws_kernel(float* __restrict__ a, const float* __restrict__ Yt, const float* __restrict__ vimg, int m, int n,
                           int j0, int ncols, int rblocks, int ntiles, int gsz) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;  // shared-window addresses throughout
  const uint32_t Bt = sbase;                                  // 4 K-chunks x {hi, lo} x (128 columns x 128 B)
  const uint32_t ring = sbase + 8 * 16384;                    // kU2Slots x 16 KB (one hi or lo image each)
  const uint32_t epi = ring + kU2Slots * 16384;               // 8 warps x 4 KB
  unsigned char* gbase = smem_dyn + (sbase - tc::smem_u32(smem_dyn));
  uint64_t* full = reinterpret_cast<uint64_t*>(gbase + 8 * 16384 + kU2Slots * 16384 + 8 * 4096);
  uint64_t* empty = full + kU2Slots;
  uint64_t* accf = empty + kU2Slots;
  uint64_t* acce = accf + 2;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(acce + 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Work items = (group of `gsz` row blocks, 128-column tile), group-major and dealt round-robin: at any
  // moment the CTAs are on the same few row groups, so the V images and A2 rows they share are found in L2.
  const int ngroups = (rblocks + gsz - 1) / gsz;
  const int nitems = ngroups * ntiles;
  if ((int)blockIdx.x >= nitems) return;  // uniform per CTA
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tslot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s2 = 0; s2 < kU2Slots; ++s2) {
      tc::mbar_init(full + s2, 1);
      tc::mbar_init(empty + s2, 1);
    }
    for (int q = 0; q < 2; ++q) {
      tc::mbar_init(accf + q, 1);
      tc::mbar_init(acce + q, 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  int blk_ctr = 0;  // row blocks this CTA has accumulated (same count in producer, issuer and epilogue)
  // A row block passes 8 images through the 4-slot ring, in the order lo(kc), hi(kc), kc = 0..3 (small
  // terms first): image i = 2 kc + h uses slot i & 3, and it is that slot's use number 2 * blk_ctr + (i >> 2).

  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int tile = item % ntiles, b0 = (item / ntiles) * gsz;
    const int b1 = min(rblocks, b0 + gsz);
    const int nblk = b1 - b0;
    const int cbase = j0 + kOB + tile * 128;
    const int cw = min(128, ncols - tile * 128);
    // Y tile (all threads): tile row = column of A2 (N index), 128 B of K per chunk.  The previous
    // item's MMAs are complete (its epilogue waited for the last accumulator before the barrier below).
    for (int idx = tid; idx < 4096; idx += kU2Threads) {
      const int q = idx & 7, c = (idx >> 3) & 127, kc = idx >> 10;
      const float4 y = c < cw ? tc::ldg128(Yt + (size_t)(tile * 128 + c) * kOB + 32 * kc + 4 * q)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      tc::split_sts(Bt + (2 * kc) * 16384, Bt + (2 * kc + 1) * 16384, c, q, y);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == 9) {
      // ---------------- producer: bulk copies of the pre-split V images ----------------
      auto prefetch_tile = [&](int bb) {  // a later row block's A2 tile into L2 (the epilogue reads it)
        if (bb < b1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int gr = j0 + bb * 128 + lane + 32 * i;
            if (gr < m) {
              const float* row = a + (size_t)gr * n + cbase;
#pragma unroll
              for (int l = 0; l < 4; ++l)
                if (cw > 32 * l) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 32 * l));
            }
          }
        }
      };
      prefetch_tile(b0 + 1);
      prefetch_tile(b0 + 2);
      for (int blk = 0; blk < nblk; ++blk) {
        if (lane == 0) {
          const int bc = blk_ctr + blk;
          const float* src = vimg + (size_t)(b0 + blk) * kVimgBlockFloats;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int kc = i >> 1, h = (i & 1) ^ 1;  // lo image first
            const int slot = i & 3, use = 2 * bc + (i >> 2);
            if (use > 0) tc::mbar_wait_parity(empty + slot, (uint32_t)((use - 1) & 1));  // MMAs of the previous use done
            const uint32_t mb = tc::smem_u32(full + slot);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             ring + slot * 16384),
                         "l"(src + kc * 8192 + h * 4096), "r"(16384), "r"(mb)
                         : "memory");
          }
        }
        prefetch_tile(b0 + blk + 3);
        __syncwarp();
      }
    } else if (warp == 8) {
      // ---------------- MMA issuer ----------------
      // The whole warp runs the loop (uniform control flow, so descriptors stay in uniform registers -- with a
      // lane-0 branch every MMA cost ~22 issue slots of ELECT/R2UR); one elected lane issues.  N = 128 per
      // MMA: an M128 x N64 x K8 tf32 MMA reads 6 KB of operands from shared memory (48 cycles at 128 B/clk)
      // for 34 cycles of math, N = 128 reads 8 KB for 68.
      for (int blk = 0; blk < nblk; ++blk) {
        const int bc = blk_ctr + blk, buf = bc & 1, ub = bc >> 1;
        if (ub > 0) tc::mbar_wait_parity(acce + buf, (uint32_t)((ub - 1) & 1));  // accumulator drained
        const uint32_t acc = tmem + 128 * buf;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kc = i >> 1, slot = i & 3;
          tc::mbar_wait_parity(full + slot, (uint32_t)((i >> 2) & 1));  // use 2 bc + (i >> 2)
          asm volatile("tcgen05.fence::after_thread_sync;");
          if (tc::elect_one()) {
            const uint64_t da = tc::umma_desc(ring + slot * 16384);
            const uint64_t dbh = tc::umma_desc(Bt + (2 * kc) * 16384), dbl = tc::umma_desc(Bt + (2 * kc + 1) * 16384);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // + 32 bytes of K per step = + 2 in the descriptor's address field
              if ((i & 1) == 0) {
                tc::mma_tf32(acc, da + 2 * ks, dbh + 2 * ks, (i > 0 || ks > 0) ? 1u : 0u);  // lo x hi
              } else {
                tc::mma_tf32(acc, da + 2 * ks, dbl + 2 * ks, 1u);  // hi x lo
                tc::mma_tf32(acc, da + 2 * ks, dbh + 2 * ks, 1u);  // hi x hi
              }
            }
            tc::umma_commit(empty + slot);            // slot free once these MMAs have read it
            if (i == 7) tc::umma_commit(accf + buf);  // accumulator complete
          }
          __syncwarp();
        }
      }
    } else if (warp < 8) {
      // ------- epilogue: warp w owns TMEM lane quarter w & 3 (32 rows) and column half w >> 2 (64 columns) -------
      const uint32_t S = epi + warp * 4096;
      const int wq = warp & 3, ch = warp >> 2;
      const int c = lane & 7, lr = lane >> 3;
      auto load_x = [&](int rb, int qc, float4 (&x)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int gr = rb + 32 * wq + 4 * i + lr, col = 64 * ch + 32 * qc + 4 * c;
          if (gr < m && col < cw) x[i] = tc::ldg128(a + (size_t)gr * n + cbase + col);
        }
      };
      auto to_smem = [&](uint32_t taddr) {  // thread = row of the tile; written so reads are row-major
        uint32_t r[32];
        tc::tmem_ld32(taddr, r);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4)
          tc::sts128(S + 16 * (lane * 8 + (g4 ^ (lane & 7))),
                     make_float4(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1]), __uint_as_float(r[4 * g4 + 2]),
                                 __uint_as_float(r[4 * g4 + 3])));
      };
      auto finish = [&](int rb, int qc, const float4 (&x)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + lr, gr = rb + 32 * wq + rr, col = 64 * ch + 32 * qc + 4 * c;
          if (gr < m && col < cw) {
            const float4 p = tc::lds128(S + 16 * (rr * 8 + (c ^ (rr & 7))));
            float4 y = x[i];
            y.x -= p.x; y.y -= p.y; y.z -= p.z; y.w -= p.w;
            tc::stg128(a + (size_t)gr * n + cbase + col, y);
          }
        }
      };
      for (int blk = 0; blk < nblk; ++blk) {
        const int rb = j0 + (b0 + blk) * 128, bc = blk_ctr + blk, buf = bc & 1, ub = bc >> 1;
        const uint32_t taddr = tmem + ((uint32_t)(32 * wq) << 16) + 128 * buf + 64 * ch;
        float4 x0[8], x1[8];
        load_x(rb, 0, x0);  // in flight while the accumulator is waited for
        load_x(rb, 1, x1);
        tc::mbar_wait_parity(accf + buf, (uint32_t)(ub & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");
        __syncwarp();  // the previous block's reads of S are done
        to_smem(taddr);
        __syncwarp();
        finish(rb, 0, x0);
        __syncwarp();
        to_smem(taddr + 32);
        asm volatile("tcgen05.fence::before_thread_sync;");
        tc::mbar_arrive(acce + buf);  // the MMAs of the block after next may overwrite this accumulator
        __syncwarp();
        finish(rb, 1, x1);
      }
    }
    blk_ctr += nblk;
    __syncthreads();  // item done: every MMA that reads the Y tile has completed (epilogue saw the last accumulator)
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// --------------------------------------------------- apply Q^T to one vector ----
// y <- H_n ... H_2 H_1 y, one 32-reflector block at a time with TWO grid barriers per block
// (instead of one per reflector).  For block V (unit lower trapezoidal, taus t):
//   pass A: every CTA accumulates G = V^T V (32 x 32) and w = V^T y over its rows,
//   two-level deterministic sum over the CTAs,
//   one warp runs the 32-step recurrence d_i = w_i, z_i = t_i d_i, w_j -= z_i G[i][j] (j > i),
//           which is exactly applying H_1 .. H_32 in order,
//   pass B: y -= V z.
// Rows are partitioned once over [0, m) (a CTA's rows above the block's diagonal just sit out),
// so y never migrates between CTAs.
constexpr int kApElems = kPB * kPB + kPB;  // G and w
__device__ __forceinline__ int ap_swz(int r) { return (r & 7) << 2; }

template <typename T>
__global__ void __launch_bounds__(kPanelThreads, 1)
    qr_apply_qt_kernel(const T* __restrict__ a, const T* __restrict__ taus, T* __restrict__ y,
                       T* __restrict__ gpart, T* __restrict__ gfull, int m, int n) {
  constexpr int nw = kPanelWarps;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Gs = reinterpret_cast<T*>(smem_raw);  // 1056: G and w
  T* zs = Gs + kApElems;                   // 32
  T* tiles = zs + kPB;                     // per warp: 32 x 32 row tile (XOR-swizzled) + 32 y values
  cgx::grid_group grid = cgx::this_grid();
  const int nb = gridDim.x, bid = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* tile = tiles + warp * (kPB * kPB + kPB);
  T* ysm = tile + kPB * kPB;
  const int per = (((m + nb - 1) / nb) + 3) & ~3;
  const int lo = bid * per < m ? bid * per : m;
  const int hi = lo + per < m ? lo + per : m;
  for (int j0 = 0; j0 < n; j0 += kPB) {
    const int nbw = min(kPB, n - j0);
    const bool cok = lane < nbw;
    const int rs = lo > j0 ? lo : j0;
    const int titers = hi > rs ? (hi - rs + 32 * nw - 1) / (32 * nw) : 0;  // block-uniform trip count
    // a warp's 32-row tile: 32 independent 128-byte row loads in flight, then shared memory
    auto load_tile = [&](int base) -> T {
      T v[kPB];
#pragma unroll
      for (int rr = 0; rr < kPB; ++rr) {
        const int r = base + rr;
        v[rr] = (r < hi && cok) ? a[(size_t)r * n + j0 + lane] : T(0);
      }
      const T yv = base + lane < hi ? y[base + lane] : T(0);
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < kPB; ++rr)
        tile[rr * kPB + (lane ^ ap_swz(rr))] = (base + rr < hi && cok) ? vmask<T>(v[rr], base + rr, j0, lane) : T(0);
      ysm[lane] = yv;
      __syncwarp();
      return yv;
    };
    // ---- pass A: lane c2 accumulates G[c1][c2] for all c1 (row entries broadcast 4 at a time)
    T g[kPB];
#pragma unroll
    for (int c1 = 0; c1 < kPB; ++c1) g[c1] = T(0);
    T wl = T(0);
    for (int it = 0; it < titers; ++it) {
      const int base = rs + (warp + it * nw) * 32;
      load_tile(base);
#pragma unroll 4
      for (int rr = 0; rr < kPB; ++rr) {
        const T x = tile[rr * kPB + (lane ^ ap_swz(rr))];
        wl = fma_(x, ysm[rr], wl);
#pragma unroll
        for (int cg = 0; cg < kPB / 4; ++cg) {
          T v4[4];
          lds4<T>(tile + rr * kPB + ((4 * cg) ^ ap_swz(rr)), v4);
#pragma unroll
          for (int e = 0; e < 4; ++e) g[4 * cg + e] = fma_(v4[e], x, g[4 * cg + e]);
        }
      }
    }
    for (int e = tid; e < kApElems; e += kPanelThreads) Gs[e] = T(0);
    __syncthreads();
    for (int w = 0; w < nw; ++w) {
      if (warp == w) {
#pragma unroll
        for (int c1 = 0; c1 < kPB; ++c1) Gs[c1 * kPB + lane] += g[c1];
        Gs[kPB * kPB + lane] += wl;
      }
      __syncthreads();
    }
    for (int e = tid; e < kApElems; e += kPanelThreads) gpart[(size_t)bid * kApElems + e] = Gs[e];
    __threadfence();
    grid.sync();
    // ---- level 2: CTA b <= 32 sums row b of [G; w] over the CTAs
    if (bid <= kPB) {
      const int e0 = bid * kPB + warp * 2;
      T s[2] = {T(0), T(0)};
      for (int b = lane; b < nb; b += 32) {
#pragma unroll
        for (int u = 0; u < 2; ++u) s[u] += __ldcg(gpart + (size_t)b * kApElems + e0 + u);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const T t = warp_sum(s[u]);
        if (lane == 0) gfull[e0 + u] = t;
      }
    }
    __threadfence();
    grid.sync();
    for (int e = tid; e < kApElems; e += kPanelThreads) Gs[e] = __ldcg(gfull + e);
    __syncthreads();
    if (warp == 0) {
      T w = Gs[kPB * kPB + lane];
      const T tl = cok ? taus[j0 + lane] : T(0);
#pragma unroll
      for (int i = 0; i < kPB; ++i) {
        const T zi = __shfl_sync(kFull, tl, i) * __shfl_sync(kFull, w, i);
        if (lane == i) zs[i] = zi;
        w = lane > i ? fma_(-zi, Gs[i * kPB + lane], w) : w;
      }
    }
    __syncthreads();
    // ---- pass B: lane = row of the tile, y_r -= sum_k v(r,k) z_k
    T z[kPB];
#pragma unroll
    for (int cg = 0; cg < kPB / 4; ++cg) lds4<T>(zs + 4 * cg, z + 4 * cg);
    for (int it = 0; it < titers; ++it) {
      const int base = rs + (warp + it * nw) * 32;
      const T yv = load_tile(base);
      T acc = T(0);
#pragma unroll
      for (int cg = 0; cg < kPB / 4; ++cg) {
        T v4[4];
        lds4<T>(tile + lane * kPB + ((4 * cg) ^ ap_swz(lane)), v4);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc = fma_(v4[e], z[4 * cg + e], acc);
      }
      if (base + lane < hi) y[base + lane] = yv - acc;
    }
    __syncthreads();  // the row -> warp map shifts with j0
  }
}

// x = R^{-1} y[0:n] for the n x n upper-triangular R stored in `a` (row-major, ld n): one CTA,
// blocked back substitution with the vector in shared memory.  Per 32-column block: warp 0 solves
// the diagonal block with shuffles, then every thread owns whole rows above it and subtracts its
// 32-term dot product (eight 128-bit loads per row, all rows of the block in flight together).
template <typename T>
__global__ void __launch_bounds__(1024) qr_rsolve_kernel(const T* __restrict__ a, const T* __restrict__ y,
                                                         T* __restrict__ x, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s = reinterpret_cast<T*>(smem_raw);  // n
  __shared__ T Rb[kPB][kPB + 1];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  constexpr int V = 16 / sizeof(T);
  using VT = typename V16K<T>::type;
  const bool al = (n % V == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0);
  for (int i = tid; i < n; i += nt) s[i] = y[i];
  const int nblk = (n + kPB - 1) / kPB;
  for (int kb = nblk - 1; kb >= 0; --kb) {
    const int k0 = kb * kPB, bw = min(kPB, n - k0);
    for (int e = tid; e < kPB * kPB; e += nt) {
      const int i = e / kPB, j = e % kPB;
      Rb[i][j] = (i < bw && j < bw) ? a[(size_t)(k0 + i) * n + k0 + j] : T(i == j);
    }
    __syncthreads();
    if (warp == 0) {
      T xi = lane < bw ? s[k0 + lane] : T(0);
#pragma unroll
      for (int j = kPB - 1; j >= 0; --j) {
        if (lane == j) xi = xi / Rb[j][j];
        const T xj = __shfl_sync(kFull, xi, j);
        if (lane < j) xi = fma_(-Rb[lane][j], xj, xi);
      }
      if (lane < bw) s[k0 + lane] = xi;
    }
    __syncthreads();
    for (int i = tid; i < k0; i += nt) {
      const T* row = a + (size_t)i * n + k0;
      T acc = T(0);
      if (al && bw == kPB) {
        VT v[kPB / V];
#pragma unroll
        for (int u = 0; u < kPB / V; ++u) v[u] = reinterpret_cast<const VT*>(row)[u];
#pragma unroll
        for (int u = 0; u < kPB / V; ++u) {
          const T* pv = reinterpret_cast<const T*>(&v[u]);
#pragma unroll
          for (int e = 0; e < V; ++e) acc = fma_(pv[e], s[k0 + u * V + e], acc);
        }
      } else {
        for (int j = 0; j < bw; ++j) acc = fma_(row[j], s[k0 + j], acc);
      }
      s[i] = s[i] - acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) x[i] = s[i];
}

// ------------------------------------------------------------------ host side ----
template <typename T>
struct QrLargePlan {
  int nb, nb_panel, rows_cta;
  size_t smem_panel, ws_bytes, wp_off, w2_off, t_off, gpart_off, gfull_off, y_off, tbig_off, yt_off, vimg_off, vimgw_off, wq_off;
  int ngroups, in_smem;
  bool ok;
};

template <typename T>
QrLargePlan<T> qr_large_plan(int m, int n) {
  QrLargePlan<T> pl{};
  pl.nb = grid_blocks();
  const int per = (((m + pl.nb - 1) / pl.nb) + 3) & ~3;
  pl.rows_cta = per;
  const size_t fixed_panel = (size_t)PanelCfg<T>::fixed_elems * sizeof(T);
  // The panel kernel runs one 512-thread CTA per SM and keeps the first RR * 16 rows of a CTA in
  // registers; the remaining rows go to shared memory only if they fit.
  pl.nb_panel = pl.nb / kGridCtasPerSm;
  const int per_panel = (((m + pl.nb_panel - 1) / pl.nb_panel) + 3) & ~3;
  const int reg_rows = PanelCfg<T>::RR * kPanelWarps;
  const size_t extra_bytes = (size_t)(per_panel > reg_rows ? per_panel - reg_rows : 0) * 33 * sizeof(T);
  pl.in_smem = fixed_panel + extra_bytes <= 215 * 1024;
  pl.smem_panel = fixed_panel + (pl.in_smem ? extra_bytes : 0);
  pl.ngroups = kW2MaxGroups;  // upper bound; the launch picks the count per panel
  size_t off = grid_part_elems();
  pl.gpart_off = off; off += (size_t)pl.nb * kPB * kPB;
  pl.gfull_off = off; off += kPB * kPB + 2 * kPB;
  pl.t_off = off; off += kPB * kPB;
  pl.wp_off = off; off += (size_t)pl.ngroups * kPB * pad4(n);
  pl.w2_off = off; off += (size_t)kPB * pad4(n);
  pl.y_off = off; off += pad4(m);
  pl.wq_off = off; off += (size_t)kWqSlices * kPB * pad4(n);  // pre-reduced row-group partials (narrow updates)
  pl.tbig_off = off; off += (size_t)kOB * kOB;        // T of an outer block (two-level blocking)
  pl.yt_off = off; off += (size_t)kOB * pad4(n);      // Y = T^T W, transposed
  off = (off + 255) & ~(size_t)255;
  pl.vimg_off = off;                                  // pre-split operand images of an outer block's V
  if (sizeof(T) == 4 && n > kOB) off += (size_t)((m + 127) / 128) * kVimgBlockFloats;
  pl.vimgw_off = off;                                 // the same V as MN-major slab images (W = V^T A2)
  if (sizeof(T) == 4 && n > kOB) off += (size_t)((m + 127) / 128) * kVimgBlockFloats;
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
  auto pk = pl.in_smem ? qr_panel_kernel<T, true> : qr_panel_kernel<T, false>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_panel));
  int occ = 0;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pk, kPanelThreads, pl.smem_panel));
  if (occ * kNumSMs < pl.nb_panel) return LXB_E_UNSUPPORTED;
  T* part = w;
  T* gpart = w + pl.gpart_off;
  T* gfull = w + pl.gfull_off;
  T* Tm = w + pl.t_off;
  T* Wp = w + pl.wp_off;
  T* W2 = w + pl.w2_off;
  // one 32-column panel at column jp
  auto run_panel = [&](int jp) -> int {
    int nbw = n - jp < kPB ? n - jp : kPB;
    int mm = m, nn = n, jj0 = jp;
    void* args[] = {&a, &taus, &Tm, &part, &gpart, &gfull, &mm, &nn, &jj0, &nbw};
    LXB_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)pk, dim3(pl.nb_panel), dim3(kPanelThreads), args,
                                             pl.smem_panel, st));
    count_launch();
    return 0;
  };
  // apply the block reflector of the panel at j0 (T in Tm) to the `ncols` columns that follow it
  auto trail32 = [&](int j0, int ncols) -> int {
    const int tiles = (ncols + kUpCols - 1) / kUpCols;
    const int wtiles = (ncols + kW2Tile - 1) / kW2Tile;
    // row groups: about two waves of CTAs at the kernel's occupancy, whatever the tile count is
    const int wocc = sizeof(T) == 4 ? 4 : 2;
    int ngroups = (2 * wocc * kNumSMs + wtiles - 1) / wtiles;
    // (the partial buffer holds kW2MaxGroups x 32 x n elements: narrow updates -- the inner updates of the
    //  two-level blocking, <= 96 columns -- may use more row groups than wide ones)
    const int gcap = (int)std::min<size_t>(4 * kNumSMs, (size_t)kW2MaxGroups * pad4(n) / (size_t)ncols);
    ngroups = ngroups < 16 ? 16 : (ngroups > gcap ? gcap : ngroups);
    {
      const size_t w_smem = (size_t)kW2Stages * kW2Rows * (kPB + kW2Tile) * sizeof(T);
      LXB_CUDA_TRY(cudaFuncSetAttribute(qr_wpartial_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w_smem));
      qr_wpartial_kernel<T><<<dim3(wtiles, ngroups), kW2Threads, w_smem, st>>>(a, Wp, m, n, j0, ncols, ngroups);
    }
    LXB_CUDA_CHECK_LAUNCH();
    if (ngroups > 64) {
      T* Wq = w + pl.wq_off;
      const int elems = kPB * ncols, per_slice = (ngroups + kWqSlices - 1) / kWqSlices;
      qr_wpresum_kernel<T><<<dim3((elems + 255) / 256, kWqSlices), 256, 0, st>>>(Wp, Wq, elems, ngroups, per_slice);
      LXB_CUDA_CHECK_LAUNCH();
      qr_wfinish_kernel<T><<<(ncols + kWfCols - 1) / kWfCols, 256, 0, st>>>(Wq, Tm, W2, ncols, kWqSlices);
    } else {
      qr_wfinish_kernel<T><<<(ncols + kWfCols - 1) / kWfCols, 256, 0, st>>>(Wp, Tm, W2, ncols, ngroups);
    }
    LXB_CUDA_CHECK_LAUNCH();
    constexpr int V = 16 / (int)sizeof(T);
    // tensor-core update is the default for fp32; LXB_QR_TC=0 selects the plain fp32 FMA kernel
    static const bool use_tc = [] { const char* e = getenv("LXB_QR_TC"); return !(e && atoi(e) == 0); }();
    if (use_tc && sizeof(T) == 4 && (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0)) {
      const int rblocks = (m - j0 + 127) / 128;
      int strips = (4 * 3 * kNumSMs + tiles - 1) / tiles;
      strips = strips < 1 ? 1 : (strips > rblocks ? rblocks : strips);
      const int per_strip = (rblocks + strips - 1) / strips;
      strips = (rblocks + per_strip - 1) / per_strip;
      LXB_CUDA_TRY(cudaFuncSetAttribute(qr_update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
      qr_update_tc_kernel<<<dim3(tiles, strips), kTcThreads, kTcSmem, st>>>(
          reinterpret_cast<float*>(a), reinterpret_cast<const float*>(W2), m, n, j0, ncols, rblocks, per_strip);
    } else if ((n % V == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0)) {
      constexpr int RT = UpCfg<T>::RT;
      const int rblocks = (m - j0 + RT - 1) / RT;
      // about four waves of CTAs (for balance); every CTA walks `per_strip` row blocks of its tile
      int strips = (4 * (sizeof(T) == 4 ? 2 : 1) * kNumSMs + tiles - 1) / tiles;
      strips = strips < 1 ? 1 : (strips > rblocks ? rblocks : strips);
      const int per_strip = (rblocks + strips - 1) / strips;
      strips = (rblocks + per_strip - 1) / per_strip;
      LXB_CUDA_TRY(cudaFuncSetAttribute(qr_update_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)UpCfg<T>::smem));
      qr_update_kernel<T><<<dim3(tiles, strips), 256, UpCfg<T>::smem, st>>>(a, W2, m, n, j0, ncols, rblocks,
                                                                            per_strip);
    } else {
      const int rblocks = (m - j0 + kUpRows - 1) / kUpRows;
      const size_t u_smem = (size_t)(kPB * kUpLd + kPB * kUpCols) * sizeof(T);
      LXB_CUDA_TRY(cudaFuncSetAttribute(qr_update_simple_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)u_smem));
      qr_update_simple_kernel<T><<<dim3(tiles, rblocks), 256, u_smem, st>>>(a, W2, m, n, j0, ncols);
    }
    LXB_CUDA_CHECK_LAUNCH();
    return 0;
  };
  // Two-level blocking (fp32, aligned): four panels per 128-column outer block, the trailing matrix is
  // updated once per outer block with K = 128 on the tensor cores (kernels above).  LXB_QR_TWOLEVEL=0
  // restores the panel-by-panel trailing updates.
  static const bool two_level_env = [] { const char* e = getenv("LXB_QR_TWOLEVEL"); return !(e && atoi(e) == 0); }();
  bool two_level = false;
  if constexpr (sizeof(T) == 4) {
    two_level = two_level_env && (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && n > kOB;
  }
  if (!two_level) {
    for (int j0 = 0; j0 < n; j0 += kPB) {
      int rc = run_panel(j0);
      if (rc) return rc;
      const int ncols = n - j0 - kPB;
      if (ncols <= 0) break;
      rc = trail32(j0, ncols);
      if (rc) return rc;
    }
    return 0;
  }
  if constexpr (sizeof(T) == 4) {
    float* af = reinterpret_cast<float*>(a);
    float* Wpf = reinterpret_cast<float*>(Wp);
    float* Tb = reinterpret_cast<float*>(w + pl.tbig_off);
    float* Yt = reinterpret_cast<float*>(w + pl.yt_off);
    LXB_CUDA_TRY(cudaFuncSetAttribute(qr_wbig_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWbigSmem));
    LXB_CUDA_TRY(cudaFuncSetAttribute(qr_wbig_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWbigSmem));
    LXB_CUDA_TRY(cudaFuncSetAttribute(qr_update128_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUp128Smem));
    LXB_CUDA_TRY(cudaFuncSetAttribute(qr_update128_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kU2Smem));
    LXB_CUDA_TRY(cudaFuncSetAttribute(qr_wbig_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kW3Smem));
    const size_t tb_smem = (size_t)2 * kOB * (kOB + 1) * sizeof(float);
    LXB_CUDA_TRY(cudaFuncSetAttribute(qr_tbig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb_smem));
    const size_t wp_cap = (size_t)pl.ngroups * kPB * pad4(n);  // elements reserved for row-group partials
    for (int j0 = 0; j0 < n; j0 += kOB) {
      const int jend = j0 + kOB < n ? j0 + kOB : n;
      for (int jp = j0; jp < jend; jp += kPB) {
        int rc = run_panel(jp);
        if (rc) return rc;
        const int inner = jend - jp - kPB;
        if (inner > 0) {
          rc = trail32(jp, inner);
          if (rc) return rc;
        }
      }
      const int ncols = n - jend;
      if (ncols <= 0) break;
      // T of the 128 reflectors: Gram matrix on the tensor cores, larft recurrence in one CTA
      int gg = (int)(wp_cap / ((size_t)kOB * kOB));
      gg = gg > 2 * kNumSMs ? 2 * kNumSMs : gg;
      qr_wbig_tc_kernel<true><<<dim3(1, gg), kTcThreads, kWbigSmem, st>>>(af, Wpf, m, n, j0, j0, kOB, gg);
      LXB_CUDA_CHECK_LAUNCH();
      qr_gsum_kernel<<<kOB * kOB / 256, 256, 0, st>>>(Wpf, Yt, gg);  // G parked in the (still unused) Y buffer
      LXB_CUDA_CHECK_LAUNCH();
      qr_tbig_kernel<<<1, kTbigThreads, tb_smem, st>>>(Yt, reinterpret_cast<const float*>(taus) + j0, Tb, 1);
      LXB_CUDA_CHECK_LAUNCH();
      // V operand images (both kernels below), W = V^T A2 (row-group partials), Y = T^T W, A2 -= V Y
      const int rblocks = (m - j0 + 127) / 128;
      float* vimg = reinterpret_cast<float*>(w + pl.vimg_off);
      float* vimgw = reinterpret_cast<float*>(w + pl.vimgw_off);
      qr_vsplit_kernel<<<dim3(4, rblocks), 256, 0, st>>>(af, vimg, vimgw, m, n, j0);
      LXB_CUDA_CHECK_LAUNCH();
      const int wt = (ncols + 127) / 128;
      static const bool w3 = [] { const char* e = getenv("LXB_QR_W3"); return !(e && atoi(e) == 0); }();
      const int cap = (int)(wp_cap / ((size_t)kOB * ncols));
      int ng;
      if (w3) {
        // one CTA per SM, about two waves; a group is a whole number of 128-row blocks
        ng = (2 * kNumSMs) / wt;
        ng = ng > cap ? cap : ng;
        ng = ng > kW2MaxGroups ? kW2MaxGroups : (ng < 1 ? 1 : ng);
        ng = ng > rblocks ? rblocks : ng;
        const int pb = (rblocks + ng - 1) / ng;
        ng = (rblocks + pb - 1) / pb;
        qr_wbig_ws_kernel<<<dim3(wt, ng), kW3Threads, kW3Smem, st>>>(af, vimgw, Wpf, m, n, j0, jend, ncols, rblocks, pb);
      } else {
        ng = (6 * kNumSMs + wt - 1) / wt;
        ng = ng > cap ? cap : ng;
        ng = ng > kW2MaxGroups ? kW2MaxGroups : (ng < 1 ? 1 : ng);
        qr_wbig_tc_kernel<false><<<dim3(wt, ng), kTcThreads, kWbigSmem, st>>>(af, Wpf, m, n, j0, jend, ncols, ng);
      }
      LXB_CUDA_CHECK_LAUNCH();
      qr_wfinish128_kernel<<<(ncols + kWf128Cols - 1) / kWf128Cols, 256, 0, st>>>(Wpf, Tb, Yt, ncols, ng);
      LXB_CUDA_CHECK_LAUNCH();
      const int ut = (ncols + 63) / 64;
      int strips = (4 * 2 * kNumSMs + ut - 1) / ut;
      strips = strips < 1 ? 1 : (strips > rblocks ? rblocks : strips);
      const int per_strip = (rblocks + strips - 1) / strips;
      strips = (rblocks + per_strip - 1) / per_strip;
      static const bool up2 = [] { const char* e = getenv("LXB_QR_UP2"); return !(e && atoi(e) == 0); }();
      if (up2) {
        // one persistent CTA per SM; 128-column tiles, groups of <= 32 row blocks sized for about a whole
        // number of rounds
        const int ut2 = (ncols + 127) / 128;
        const int64_t cells = (int64_t)ut2 * rblocks;
        const int rounds = (int)((cells + (int64_t)kNumSMs * 32 - 1) / ((int64_t)kNumSMs * 32));
        // row groups per tile so that groups x tiles fits `rounds` full rounds of the CTAs (rounding the group
        // SIZE instead could spill a few items into one more round: 310 items on 148 CTAs at m = 32 768)
        int gpt = (int)(((int64_t)kNumSMs * rounds) / ut2);
        gpt = gpt < 1 ? 1 : (gpt > rblocks ? rblocks : gpt);
        const int gsz = (rblocks + gpt - 1) / gpt;
        const int64_t nitems = (int64_t)((rblocks + gsz - 1) / gsz) * ut2;
        const int g2 = (int)(nitems < kNumSMs ? nitems : kNumSMs);
        qr_update128_ws_kernel<<<g2, kU2Threads, kU2Smem, st>>>(af, Yt, vimg, m, n, j0, ncols, rblocks, ut2, gsz);
      } else {
        qr_update128_tc_kernel<<<dim3(ut, strips), kTcThreads, kUp128Smem, st>>>(af, Yt, m, n, j0, ncols, rblocks, per_strip);
      }
      LXB_CUDA_CHECK_LAUNCH();
    }
  }
  return 0;
}

// least squares with the large factors: x = R^{-1} (Q^T b)[:n]
template <typename T>
int qr_large_solve(const T* a, const T* taus, const T* b, T* x, int m, int n, void* ws, size_t ws_bytes,
                   cudaStream_t st, bool qt_only) {
  const QrLargePlan<T> pl = qr_large_plan<T>(m, n);
  if (!pl.ok) return LXB_E_UNSUPPORTED;
  if (!ws || ws_bytes < pl.ws_bytes) return LXB_E_WORKSPACE;
  T* w = reinterpret_cast<T*>(ws);
  T* y = w + pl.y_off;
  LXB_CUDA_TRY(cudaMemcpyAsync(y, b, (size_t)m * sizeof(T), cudaMemcpyDeviceToDevice, st));
  count_launch();
  auto ak = qr_apply_qt_kernel<T>;
  T* gpart = w + pl.gpart_off;  // nb x 1024 elements reserved, 148 x 1056 used
  T* gfull = w + pl.gfull_off;
  int mm = m, nn = n;
  void* args[] = {&a, &taus, &y, &gpart, &gfull, &mm, &nn};
  const size_t ap_smem = (size_t)(kApElems + kPB + kPanelWarps * (kPB * kPB + kPB)) * sizeof(T);
  LXB_CUDA_TRY(cudaFuncSetAttribute(ak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ap_smem));
  LXB_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)ak, dim3(pl.nb_panel), dim3(kPanelThreads), args, ap_smem,
                                           st));
  count_launch();
  if (qt_only) {  // (Q^T b)[:n] only: the row-sharded TSQR stacks these over the ranks
    LXB_CUDA_TRY(cudaMemcpyAsync(x, y, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    count_launch();
    return 0;
  }
  const size_t smem = (size_t)n * sizeof(T);
  if (smem > 200 * 1024) return LXB_E_UNSUPPORTED;
  LXB_CUDA_TRY(cudaFuncSetAttribute(qr_rsolve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qr_rsolve_kernel<T><<<1, 1024, smem, st>>>(a, y, x, n);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
size_t qr_large_ws_bytes(int m, int n) { return qr_large_plan<T>(m, n).ws_bytes; }

#ifdef LXB_QR_PROF
}  // namespace lxb
extern "C" void lxb_debug_qr_prof(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, lxb::g_prof, sizeof(lxb::g_prof));
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(lxb::g_prof, z, sizeof(z));
}
namespace lxb {
#endif

#define LXB_INST_QRL(T)                                                                             \
  template int qr_large_factor<T>(const T*, T*, T*, int, int, void*, size_t, cudaStream_t);         \
  template int qr_large_solve<T>(const T*, const T*, const T*, T*, int, int, void*, size_t, cudaStream_t, bool); \
  template size_t qr_large_ws_bytes<T>(int, int);
LXB_INST_QRL(float)
LXB_INST_QRL(double)

}  // namespace lxb

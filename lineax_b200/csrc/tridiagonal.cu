// Batched tridiagonal solve, lineax/_solver/tridiagonal.py:54-72
// (lax.linalg.tridiagonal_solve -> LAPACK gtsv on CPU, cuSPARSE gtsv2 on GPU).
//
// Two kernels:
//  * tridiag_warp_kernel (below, second half of the file): ONE WARP per system, the system resident in
//    registers (R consecutive rows per lane), for row-diagonally-dominant systems -- the class for which
//    elimination without row interchanges is backward stable, so the result agrees with gtsv's to
//    rounding.  Every lane eliminates its R-1 interior rows locally (Thomas with two spike right-hand
//    sides), the 32 separator rows form a reduced tridiagonal system solved by parallel cyclic reduction
//    over the lanes with shuffles, and the interior unknowns follow with two FMAs per row.  Nothing
//    but the four operands (in) and x (out) crosses HBM: 5 n sizeof(T) bytes per system.
//    Operands are staged with 16-byte cp.async into an XOR-swizzled shared-memory row per array (the
//    511-element off-diagonals are copied as the 16-byte-aligned superset and shifted in registers),
//    one system ahead of the arithmetic.  A system that is not diagonally dominant is appended to a
//    list instead, and
//  * tridiagonal_kernel (first half): one THREAD per system running gtsv's own pivoting elimination,
//    handles the listed systems (and every shape the warp kernel does not cover).
//
// tridiagonal_kernel
// Design: one THREAD per system running gtsv's own elimination (row interchange when
// |d_i| < |dl_i|, second super-diagonal fill-in), so results follow LAPACK for every
// input, not just diagonally dominant ones.  HBM access stays coalesced because each
// warp moves its 32 systems through shared memory in 16-element chunks: cp.async copies
// (64-byte row segments in, transposed conflict-free reads out) run one chunk AHEAD of the
// serial recurrence through a two-stage ring, so the recurrence never waits on a load.
// The forward sweep leaves the normalised rows (u1 = du/d, u2 = du2/d) in a per-warp scratch
// slab and y = b/d in the output buffer; the backward sweep streams them back the same way.
// Algorithmic HBM traffic: 5 n sizeof(T) per system (4 arrays in, x out).
#include <stdlib.h>

#include "common.cuh"

namespace lxb {

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

constexpr int kTriWarps = 4;  // warps per CTA
constexpr int kCh = 16;       // elements per chunk
constexpr int kTileLd = 33;   // [element][system] tiles, padded

__device__ __forceinline__ void tri_cp(void* smem_dst, const void* gmem_src, int bytes, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sb = ok ? bytes : 0;
  if (bytes == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sb));
  else if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sb));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sb));
}

template <typename T>
__global__ void __launch_bounds__(kTriWarps * 32)
    tridiagonal_kernel(const T* __restrict__ D, const T* __restrict__ DL, const T* __restrict__ DU,
                       int64_t sD, const T* __restrict__ B, int64_t sB, T* __restrict__ X,
                       T* __restrict__ ws, int64_t batch, int n, const int32_t* __restrict__ list,
                       const int32_t* __restrict__ list_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kTileElems = kCh * kTileLd;
  constexpr int kStage = 4 * kTileElems;  // forward: d, dl, du, b; backward: y, u1, u2 (fits)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* wsm = reinterpret_cast<T*>(smem_raw) + (size_t)warp * 2 * kStage;
  const int64_t wslot = (int64_t)blockIdx.x * kTriWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kTriWarps;
  T* w1 = ws + wslot * 2 * (int64_t)n * 32;  // u1[i][lane]
  T* w2 = w1 + (int64_t)n * 32;              // u2[i][lane]
  const int64_t sOff = sD ? sD - 1 : 0;      // off-diagonals hold n-1 entries per system
  if (list != nullptr) batch = *list_count;  // only the systems the warp kernel handed over
  const int64_t groups = (batch + 31) / 32;
  // system handled by row r of group g (identity, or an entry of the hand-over list)
  auto sysid = [&](int64_t g, int r) -> int64_t { return list ? (int64_t)list[g * 32 + r] : g * 32 + r; };
  const int nchunks = (n + kCh - 1) / kCh;
  const int half = lane >> 4, el = lane & 15;  // copy mapping: two systems x 16 elements per step

  for (int64_t g = wslot; g < groups; g += nwarps) {
    const int64_t sys0 = g * 32;
    const int nsys = (int)((batch - sys0) < 32 ? (batch - sys0) : 32);
    auto issue_fwd = [&](int st, int c) {
      T* td = wsm + st * kStage;
      T* tl = td + kTileElems;
      T* tu = tl + kTileElems;
      T* tb = tu + kTileElems;
      const int i = c * kCh + el;
#pragma unroll 4
      for (int rp = 0; rp < 16; ++rp) {
        const int r = 2 * rp + half;
        const bool ok = r < nsys && i < n;
        const int64_t s = ok ? sysid(g, r) : 0;
        const int o = el * kTileLd + r;
        tri_cp(td + o, ok ? D + s * sD + i : D, (int)sizeof(T), ok);
        const bool okl = ok && i >= 1;  // sub-diagonal entry of row i
        tri_cp(tl + o, okl ? DL + s * sOff + i - 1 : D, (int)sizeof(T), okl);
        const bool oku = ok && i < n - 1;
        tri_cp(tu + o, oku ? DU + s * sOff + i : D, (int)sizeof(T), oku);
        tri_cp(tb + o, ok ? B + s * sB + i : D, (int)sizeof(T), ok);
      }
    };
    // carried state of the current row (row i), initialised at the first chunk
    T cd = T(0), cu = T(0), cb = T(0);
    // -------- forward sweep
    __syncwarp();
    issue_fwd(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
      const int i0 = c * kCh;
      __syncwarp();  // everyone is done reading the stage that is refilled next
      if (c + 1 < nchunks) issue_fwd((c + 1) & 1, c + 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
      T* td = wsm + (c & 1) * kStage;
      T* tl = td + kTileElems;
      T* tu = tl + kTileElems;
      T* tb = tu + kTileElems;
      const int cnt = (n - i0) < kCh ? (n - i0) : kCh;
      if (lane < nsys) {
        for (int e = 0; e < cnt; ++e) {
          const int i = i0 + e;
          if (i == 0) {
            cd = td[lane];
            cu = tu[lane];
            cb = tb[lane];
            continue;
          }
          // eliminate the sub-diagonal entry of row i against the carried row i-1
          const T dl = tl[e * kTileLd + lane], dn = td[e * kTileLd + lane], un = tu[e * kTileLd + lane],
                  bn = tb[e * kTileLd + lane];
          // gtsv's step with ONE division per row: the pivot of the finished row (od) is inverted
          // once and both the multiplier (fact = other / od) and the normalised row use it
          T ou, ou2, ob, rinv;  // finished row i-1 (scaled by 1/od below)
          if (abs_(cd) >= abs_(dl)) {
            rinv = T(1) / cd;
            const T fact = dl * rinv;
            ou = cu; ou2 = T(0); ob = cb;
            cd = dn - fact * cu;
            cb = bn - fact * cb;
            cu = un;
          } else {
            rinv = T(1) / dl;
            const T fact = cd * rinv;
            ou = dn; ou2 = un; ob = bn;
            cd = cu - fact * dn;
            cb = cb - fact * bn;
            cu = -fact * un;
          }
          w1[(int64_t)(i - 1) * 32 + lane] = ou * rinv;
          w2[(int64_t)(i - 1) * 32 + lane] = ou2 * rinv;
          // y_{i-1} goes to the output tile slot of element e-1; the previous chunk's last element
          // is handed over through td's first slot (see the store below)
          if (e > 0) tb[(e - 1) * kTileLd + lane] = ob * rinv;
          else td[lane] = ob * rinv;
        }
        if (i0 + cnt == n) tb[(cnt - 1) * kTileLd + lane] = cb / cd;  // last row: y_{n-1} = x_{n-1}
      }
      __syncwarp();
      // coalesced store of y for this chunk (elements i0 .. i0+cnt-1, except the chunk's last
      // element which is only known after the next chunk's first step) and of element i0-1
      const bool last_chunk = i0 + kCh >= n;
      const int upto = last_chunk ? n : i0 + kCh - 1;  // exclusive bound of final values
#pragma unroll 4
      for (int rp = 0; rp < 16; ++rp) {
        const int r = 2 * rp + half;
        const int i = i0 + el;
        if (r < nsys && i < upto) X[sysid(g, r) * n + i] = tb[el * kTileLd + r];
      }
      if (i0 > 0 && lane < nsys) X[sysid(g, lane) * n + i0 - 1] = td[lane];
    }
    // -------- backward sweep: x_i = y_i - u1_i x_{i+1} - u2_i x_{i+2}
    auto issue_bwd = [&](int st, int c) {
      T* ty = wsm + st * kStage;
      T* t1 = ty + kTileElems;      // [kCh][32]
      T* t2 = t1 + kCh * 32;        // [kCh][32]
      const int i0 = c * kCh;
      const int i = i0 + el;
#pragma unroll 4
      for (int rp = 0; rp < 16; ++rp) {
        const int r = 2 * rp + half;
        const bool ok = r < nsys && i < n;
        tri_cp(ty + el * kTileLd + r, ok ? X + sysid(g, r) * n + i : X, (int)sizeof(T), ok);
      }
      constexpr int V = 16 / (int)sizeof(T);  // elements per 16-byte piece
      constexpr int ppr = 32 / V;             // pieces per scratch row
#pragma unroll
      for (int q = 0; q < kCh * ppr / 32; ++q) {
        const int pc = lane + q * 32, e = pc / ppr, off = (pc % ppr) * V;
        const bool ok = i0 + e < n - 1;
        tri_cp(t1 + e * 32 + off, ok ? w1 + (int64_t)(i0 + e) * 32 + off : w1, 16, ok);
        tri_cp(t2 + e * 32 + off, ok ? w2 + (int64_t)(i0 + e) * 32 + off : w2, 16, ok);
      }
    };
    T x1 = T(0), x2 = T(0);  // x_{i+1}, x_{i+2}
    __syncwarp();
    issue_bwd(0, nchunks - 1);
    cp_async_commit();
    for (int k = 0; k < nchunks; ++k) {
      const int c = nchunks - 1 - k;
      const int i0 = c * kCh;
      __syncwarp();
      if (c > 0) issue_bwd((k + 1) & 1, c - 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
      T* ty = wsm + (k & 1) * kStage;
      const T* t1 = ty + kTileElems;
      const T* t2 = t1 + kCh * 32;
      const int cnt = (n - i0) < kCh ? (n - i0) : kCh;
      if (lane < nsys) {
        for (int e = cnt - 1; e >= 0; --e) {
          const int i = i0 + e;
          T xi = ty[e * kTileLd + lane];
          if (i < n - 1) xi = xi - t1[e * 32 + lane] * x1 - t2[e * 32 + lane] * x2;
          ty[e * kTileLd + lane] = xi;
          x2 = x1;
          x1 = xi;
        }
      }
      __syncwarp();
#pragma unroll 4
      for (int rp = 0; rp < 16; ++rp) {
        const int r = 2 * rp + half;
        const int i = i0 + el;
        if (r < nsys && i < n) X[sysid(g, r) * n + i] = ty[el * kTileLd + r];
      }
    }
    cp_async_wait<0>();
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-per-system kernel for diagonally dominant systems (see the header comment).
constexpr int kTwWarps = 4;  // warps per CTA

template <typename T>
__device__ __forceinline__ T tri_rcp(T x) {
  if constexpr (sizeof(T) == 4) {
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(x));
    return fma_(r0, fma_(-x, r0, 1.0f), r0);  // MUFU.RCP + one Newton step (correctly rounded for normal x)
  } else {
    return T(1) / x;
  }
}

// chunks (16 bytes) per staged array row: n elements plus the alignment shift, rounded up
template <typename T>
__host__ __device__ constexpr int tw_row_chunks(int R) {
  return R * 32 / (16 / (int)sizeof(T)) + 4;
}
// physical position of logical chunk q: XOR of the low bits with the bits that select the lane group, so
// that the 8 lanes served by one LDS.128 phase (chunks CPL apart) hit 8 different 16-byte bank groups
template <int CPL>
__device__ __forceinline__ int tw_swz(int q) {
  return q ^ ((q >> 3) & (CPL - 1));
}

// Stage elements [e0, e0 + n) of `base` (element offsets; the array holds `total` elements) into the
// swizzled row at `dst`: the 16-byte-aligned superset is copied, row i lands at element t + i where
// t = (absolute element index of e0) mod V is returned by tw_shift().
template <typename T>
__device__ __forceinline__ int tw_shift(const T* base, int64_t e0) {
  constexpr int V = 16 / (int)sizeof(T);
  const int64_t a0 = (int64_t)((reinterpret_cast<uintptr_t>(base) / sizeof(T)) % V) + e0;
  return (int)(((a0 % V) + V) % V);
}
template <typename T, int R>
__device__ __forceinline__ void tw_stage(T* dst, const T* base, int64_t e0, int n, int64_t total, int lane) {
  constexpr int V = 16 / (int)sizeof(T);
  constexpr int CPL = R / V;
  constexpr int KMAX = (R * 32 / V + 1 + 31) / 32;  // chunks per lane: n / V + 1 at most
  const int t = tw_shift(base, e0);
  const int nch = (t + n + V - 1) / V;
  // one 64-bit pointer per lane; every chunk of the lane is a constant 512 bytes further on
  const char* src = reinterpret_cast<const char*>(base + (e0 - t)) + lane * 16;
  const unsigned dsts = (unsigned)__cvta_generic_to_shared(dst);
  // any aligned 16-byte piece that overlaps the array is safe to read in full (allocation granularity);
  // pieces entirely outside (before element 0 / at or after element `total`) are zero-filled unread
  const int64_t rem64 = total - (e0 - t);
  const int rem = rem64 > (int64_t)(1 << 30) ? (1 << 30) : (int)rem64;  // elements from chunk 0 to the end
  const bool first_ok = e0 - t + V > 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int q = lane + 32 * k;
    if (q < nch) {
      const bool ok = q * V < rem && (q > 0 || first_ok);
      const int sb = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dsts + tw_swz<CPL>(q) * 16),
                   "l"(ok ? (const void*)(src + k * 512) : (const void*)base), "r"(sb));
    }
  }
}
// Lane's R consecutive rows (starting at row R * lane) of a staged array, undoing the alignment shift t.
template <typename T, int R>
__device__ __forceinline__ void tw_read(const T* row, int t, int lane, T (&v)[R]) {
  constexpr int V = 16 / (int)sizeof(T);
  constexpr int CPL = R / V;
  using VT = typename V16<T>::type;
  T buf[R + V];
#pragma unroll
  for (int k = 0; k <= CPL; ++k) {
    const VT c = *reinterpret_cast<const VT*>(reinterpret_cast<const char*>(row) + tw_swz<CPL>(CPL * lane + k) * 16);
    const T* pc = reinterpret_cast<const T*>(&c);
#pragma unroll
    for (int e = 0; e < V; ++e) buf[k * V + e] = pc[e];
  }
  // t is warp-uniform: two conditional shifts (by 1 and by 2 elements) instead of a 4-way unrolled switch
  if (t & 1) {
#pragma unroll
    for (int j = 0; j < R + V - 1; ++j) buf[j] = buf[j + 1];
  }
  if (V > 2 && (t & 2)) {
#pragma unroll
    for (int j = 0; j < R + V - 2; ++j) buf[j] = buf[j + 2];
  }
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = buf[j];
}

template <typename T, int R>
__global__ void __launch_bounds__(kTwWarps * 32)
    tridiag_warp_kernel(const T* __restrict__ D, const T* __restrict__ DL, const T* __restrict__ DU, int64_t sD,
                        const T* __restrict__ B, int64_t sB, T* __restrict__ X, int64_t batch, int n,
                        int32_t* __restrict__ list, int32_t* __restrict__ list_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int V = 16 / (int)sizeof(T);
  constexpr int CPL = R / V;
  constexpr int RC = tw_row_chunks<T>(R);   // chunks per staged array
  constexpr int STAGE = 4 * RC * 16;        // bytes per stage: d, dl, du, b
  using VT = typename V16<T>::type;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wsm = smem_raw + (size_t)warp * 2 * STAGE;
  const int64_t nwarps = (int64_t)gridDim.x * kTwWarps;
  const int64_t sOff = sD ? sD - 1 : 0;  // off-diagonals hold n-1 entries per system
  const int64_t totD = (batch - 1) * sD + n, totO = (batch - 1) * sOff + (n - 1), totB = (batch - 1) * sB + n;
  auto stage_sys = [&](int st, int64_t s) {
    T* r0 = reinterpret_cast<T*>(wsm + st * STAGE);
    tw_stage<T, R>(r0, D, s * sD, n, totD, lane);
    tw_stage<T, R>(r0 + RC * V, DL, s * sOff - 1, n, totO, lane);      // row i <-> DL[s*sOff + i - 1]
    tw_stage<T, R>(r0 + 2 * RC * V, DU, s * sOff, n, totO, lane);      // row i <-> DU[s*sOff + i]
    tw_stage<T, R>(r0 + 3 * RC * V, B, s * sB, n, totB, lane);
    cp_async_commit();
  };
  int64_t sys = (int64_t)blockIdx.x * kTwWarps + warp;
  if (sys < batch) stage_sys(0, sys);
  int st = 0;
  for (; sys < batch; sys += nwarps, st ^= 1) {
    const int64_t nxt = sys + nwarps;
    __syncwarp();  // the other stage (x staging of the previous system) has been drained
    if (nxt < batch) stage_sys(st ^ 1, nxt);
    else cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    T* r0 = reinterpret_cast<T*>(wsm + st * STAGE);
    T a[R], d[R], c[R], b[R];
    tw_read<T, R>(r0, tw_shift(D, sys * sD), lane, d);
    tw_read<T, R>(r0 + RC * V, tw_shift(DL, sys * sOff - 1), lane, a);
    tw_read<T, R>(r0 + 2 * RC * V, tw_shift(DU, sys * sOff), lane, c);
    tw_read<T, R>(r0 + 3 * RC * V, tw_shift(B, sys * sB), lane, b);
    // rows outside the system: identity rows; first / last row have no outer coupling
    if (n == 32 * R) {  // (warp-uniform) full tile: only two entries to clear
      if (lane == 0) a[0] = T(0);
      if (lane == 31) c[R - 1] = T(0);
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int i = R * lane + j;
        if (i >= n) { d[j] = T(1); b[j] = T(0); }
        if (i >= n || i == 0) a[j] = T(0);
        if (i >= n - 1) c[j] = T(0);
      }
    }
    bool dominant = true;
#pragma unroll
    for (int j = 0; j < R; ++j) dominant = dominant && (abs_(d[j]) >= abs_(a[j]) + abs_(c[j])) && d[j] != T(0);
    if (!__all_sync(kFull, dominant)) {
      // not diagonally dominant (or NaN): hand the system to the pivoting kernel
      if (lane == 0) list[atomicAdd(list_count, 1)] = (int32_t)sys;
      continue;
    }
    constexpr int M = R - 1;  // interior rows 0..M-1, separator row M
    // ---- local elimination of the interior rows: d <- 1/pivot, b <- g, a <- W (spike of the coupling to
    //      the previous separator), c <- V (spike of the coupling to this lane's separator)
    d[0] = tri_rcp(d[0]);
#pragma unroll
    for (int j = 1; j < M; ++j) {
      const T m = a[j] * d[j - 1];
      d[j] = tri_rcp(fma_(-m, c[j - 1], d[j]));
      b[j] = fma_(-m, b[j - 1], b[j]);
      a[j] = -m * a[j - 1];
    }
    if (M >= 1) {
      b[M - 1] *= d[M - 1];
      a[M - 1] *= d[M - 1];
      c[M - 1] *= d[M - 1];
#pragma unroll
      for (int j = M - 2; j >= 0; --j) {
        const T t = c[j] * d[j];
        b[j] = fma_(-t, b[j + 1], b[j] * d[j]);
        a[j] = fma_(-t, a[j + 1], a[j] * d[j]);
        c[j] = -t * c[j + 1];
      }
    }
    // ---- reduced system on the 32 separators: al s_{l-1} + be s_l + ga s_{l+1} = de
    const T g0n = __shfl_down_sync(kFull, b[0], 1), w0n = __shfl_down_sync(kFull, a[0], 1),
            v0n = __shfl_down_sync(kFull, c[0], 1);
    T al, be, ga, de;
    if (M >= 1) {
      al = -a[M] * a[M - 1];
      be = fma_(-c[M], w0n, fma_(-a[M], c[M - 1], d[M]));
      ga = -c[M] * v0n;
      de = fma_(-c[M], g0n, fma_(-a[M], b[M - 1], b[M]));
    } else {
      al = a[M]; be = d[M]; ga = c[M]; de = b[M];
    }
    if (lane == 31) ga = T(0);
    if (lane == 0) al = T(0);
    // parallel cyclic reduction over the lanes (out-of-range neighbours enter with coupling 0)
#pragma unroll
    for (int h = 1; h < 32; h <<= 1) {
      const T alm = __shfl_up_sync(kFull, al, h), bem = __shfl_up_sync(kFull, be, h),
              gam = __shfl_up_sync(kFull, ga, h), dem = __shfl_up_sync(kFull, de, h);
      const T alp = __shfl_down_sync(kFull, al, h), bep = __shfl_down_sync(kFull, be, h),
              gap = __shfl_down_sync(kFull, ga, h), dep = __shfl_down_sync(kFull, de, h);
      const T k1 = lane >= h ? al * tri_rcp(bem) : T(0);
      const T k2 = lane + h < 32 ? ga * tri_rcp(bep) : T(0);
      be = fma_(-k2, alp, fma_(-k1, gam, be));
      de = fma_(-k2, dep, fma_(-k1, dem, de));
      al = -k1 * alm;
      ga = -k2 * gap;
    }
    const T sep = de * tri_rcp(be);
    T sprev = __shfl_up_sync(kFull, sep, 1);
    if (lane == 0) sprev = T(0);
    // ---- interior unknowns and coalesced store through the (drained) stage
#pragma unroll
    for (int j = 0; j < M; ++j) b[j] = fma_(-c[j], sep, fma_(-a[j], sprev, b[j]));
    b[M] = sep;
    __syncwarp();  // every lane has read its operands: the d row of this stage can take x
    {
      char* xr = reinterpret_cast<char*>(r0);
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        VT v;
        T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
        for (int e = 0; e < V; ++e) pv[e] = b[k * V + e];
        *reinterpret_cast<VT*>(xr + tw_swz<CPL>(CPL * lane + k) * 16) = v;
      }
      __syncwarp();
      VT* xo = reinterpret_cast<VT*>(X + sys * n);
      const int nch = n / V;
      for (int q = lane; q < nch; q += 32) xo[q] = *reinterpret_cast<const VT*>(xr + tw_swz<CPL>(q) * 16);
    }
  }
  cp_async_wait<0>();
}

template <typename T>
struct TriPlan {
  int blocks;
  size_t smem, scratch_bytes, ws_bytes;
};

template <typename T>
TriPlan<T> tri_plan(int64_t batch, int n) {
  TriPlan<T> pl{};
  pl.smem = (size_t)kTriWarps * 2 * 4 * kCh * kTileLd * sizeof(T);
  const int64_t groups = (batch + 31) / 32;
  int64_t blocks = (groups + kTriWarps - 1) / kTriWarps;
  // scratch slab: 2 n 32 sizeof(T) per warp.  Measured on B200 (2^20 x 512 f32): latency hiding
  // needs every resident warp (12 per SM), which matters more than keeping the slab inside L2.
  const size_t per_block = (size_t)kTriWarps * 2 * (size_t)n * 32 * sizeof(T);
  const int64_t cap = (int64_t)kNumSMs * (sizeof(T) == 4 ? 3 : 1);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  pl.blocks = (int)blocks;
  pl.scratch_bytes = (size_t)blocks * per_block;
  // + hand-over list of the warp kernel: counter (16 bytes) and one int32 per system
  pl.ws_bytes = pl.scratch_bytes + 16 + (((size_t)batch * 4 + 15) & ~(size_t)15);
  return pl;
}

template <typename T, int R>
int launch_tridiag_warp(const T* d, const T* dl, const T* du, int64_t sD, const T* b, int64_t sb, T* x,
                        int64_t batch, int n, int32_t* list, int32_t* count, cudaStream_t st) {
  auto kern = tridiag_warp_kernel<T, R>;
  const size_t smem = (size_t)kTwWarps * 2 * 4 * tw_row_chunks<T>(R) * 16;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTwWarps * 32, smem));
  if (occ < 1) occ = 1;
  int64_t blocks = (batch + kTwWarps - 1) / kTwWarps;
  if (blocks > (int64_t)kNumSMs * occ) blocks = (int64_t)kNumSMs * occ;
  kern<<<(unsigned)blocks, kTwWarps * 32, smem, st>>>(d, dl, du, sD, b, sb, x, batch, n, list, count);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int tridiagonal_solve(const T* d, const T* dl, const T* du, int64_t sD, const T* b, int64_t sb, T* x,
                      int64_t batch, int n, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (batch < 0 || n < 0 || !d || !b || !x || (n > 1 && (!dl || !du))) return LXB_E_BADARG;
  if (batch == 0 || n == 0) return 0;
  const TriPlan<T> pl = tri_plan<T>(batch, n);
  if (!ws || ws_bytes < pl.ws_bytes) return LXB_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 15) return LXB_E_ALIGN;
  auto kern = tridiagonal_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  constexpr int V = 16 / (int)sizeof(T);
  constexpr int kMaxR = sizeof(T) == 4 ? 32 : 16;  // rows per lane the register budget allows
  // warp-per-system kernel: long enough systems with 16-byte aligned rows of x; stride-0 (broadcast)
  // operands and everything else stay on the thread-per-system kernel
  static const bool no_warp = getenv("LXB_TRIDIAG_LEGACY") != nullptr;
  const bool fast = !no_warp && n >= 64 && n <= 32 * kMaxR && n % V == 0 && aligned16(x) && sD >= n && sb >= n &&
                    batch < (int64_t)1 << 31;
  if (!fast) {
    kern<<<pl.blocks, kTriWarps * 32, pl.smem, st>>>(d, dl, du, sD, b, sb, x, reinterpret_cast<T*>(ws), batch, n,
                                                     nullptr, nullptr);
    LXB_CUDA_CHECK_LAUNCH();
    return 0;
  }
  int32_t* count = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + pl.scratch_bytes);
  int32_t* list = count + 4;
  LXB_CUDA_TRY(cudaMemsetAsync(count, 0, 16, st));
  int rc;
  if (n <= 32 * 4) rc = launch_tridiag_warp<T, 4>(d, dl, du, sD, b, sb, x, batch, n, list, count, st);
  else if (n <= 32 * 8) rc = launch_tridiag_warp<T, 8>(d, dl, du, sD, b, sb, x, batch, n, list, count, st);
  else if (n <= 32 * 16) rc = launch_tridiag_warp<T, 16>(d, dl, du, sD, b, sb, x, batch, n, list, count, st);
  else rc = launch_tridiag_warp<T, kMaxR>(d, dl, du, sD, b, sb, x, batch, n, list, count, st);
  if (rc != 0) return rc;
  // systems that were not diagonally dominant (none for the BASELINE generator): pivoting elimination
  kern<<<pl.blocks, kTriWarps * 32, pl.smem, st>>>(d, dl, du, sD, b, sb, x, reinterpret_cast<T*>(ws), batch, n,
                                                   list, count);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_TRIDIAG(sfx, T)                                                                    \
  extern "C" int lxb_tridiagonal_solve_##sfx(const T* d, const T* dl, const T* du,                 \
                                             int64_t stride_diag, const T* b, int64_t stride_b,    \
                                             T* x, int64_t batch, int32_t n, void* workspace,      \
                                             size_t workspace_bytes, lxb_stream_t stream) {        \
    return lxb::tridiagonal_solve<T>(d, dl, du, stride_diag, b, stride_b, x, batch, n, workspace,  \
                                     workspace_bytes, (cudaStream_t)stream);                       \
  }                                                                                                \
  extern "C" size_t lxb_tridiagonal_workspace_##sfx(int64_t batch, int32_t n) {                    \
    if (batch <= 0 || n <= 0) return 0;                                                            \
    return lxb::tri_plan<T>(batch, n).ws_bytes;                                                    \
  }
LXB_DEF_TRIDIAG(f32, float)
LXB_DEF_TRIDIAG(f64, double)

// Batched tridiagonal solve with partial pivoting, lineax/_solver/tridiagonal.py:54-72
// (lax.linalg.tridiagonal_solve -> LAPACK gtsv on CPU, cuSPARSE gtsv2 on GPU).
//
// Design: one THREAD per system running gtsv's own elimination (row interchange when
// |d_i| < |dl_i|, second super-diagonal fill-in), so results follow LAPACK for every
// input, not just diagonally dominant ones.  HBM access stays coalesced because each
// warp moves its 32 systems through shared memory in 16-element chunks: cp.async copies
// (64-byte row segments in, transposed conflict-free reads out) run one chunk AHEAD of the
// serial recurrence through a two-stage ring, so the recurrence never waits on a load.
// The forward sweep leaves the normalised rows (u1 = du/d, u2 = du2/d) in a per-warp scratch
// slab and y = b/d in the output buffer; the backward sweep streams them back the same way.
// Algorithmic HBM traffic: 5 n sizeof(T) per system (4 arrays in, x out).
#include "common.cuh"

namespace lxb {

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
                       T* __restrict__ ws, int64_t batch, int n) {
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
  const int64_t groups = (batch + 31) / 32;
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
        const int64_t s = sys0 + r;
        const bool ok = r < nsys && i < n;
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
        if (r < nsys && i < upto) X[(sys0 + r) * n + i] = tb[el * kTileLd + r];
      }
      if (i0 > 0 && lane < nsys) X[(sys0 + lane) * n + i0 - 1] = td[lane];
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
        tri_cp(ty + el * kTileLd + r, ok ? X + (sys0 + r) * n + i : X, (int)sizeof(T), ok);
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
        if (r < nsys && i < n) X[(sys0 + r) * n + i] = ty[el * kTileLd + r];
      }
    }
    cp_async_wait<0>();
  }
}

template <typename T>
struct TriPlan {
  int blocks;
  size_t smem, ws_bytes;
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
  pl.ws_bytes = (size_t)blocks * per_block;
  return pl;
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
  kern<<<pl.blocks, kTriWarps * 32, pl.smem, st>>>(d, dl, du, sD, b, sb, x, reinterpret_cast<T*>(ws),
                                                   batch, n);
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

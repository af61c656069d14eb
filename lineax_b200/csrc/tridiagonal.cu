// Batched tridiagonal solve with partial pivoting, lineax/_solver/tridiagonal.py:54-72
// (lax.linalg.tridiagonal_solve -> LAPACK gtsv on CPU, cuSPARSE gtsv2 on GPU).
//
// Design: one THREAD per system running gtsv's own elimination (row interchange when
// |d_i| < |dl_i|, second super-diagonal fill-in), so results follow LAPACK for every
// input, not just diagonally dominant ones.  HBM access stays coalesced because each
// warp moves its 32 systems through shared memory in 32x32-element tiles (128-byte row
// segments in, transposed conflict-free reads out).  The forward sweep leaves the
// normalised rows (u1 = du/d, u2 = du2/d) in a per-warp scratch slab that is sized to
// stay L2-resident and y = b/d in the output buffer; the backward sweep streams them back.
// Algorithmic HBM traffic: 5 n sizeof(T) per system (4 arrays in, x out).
#include "common.cuh"

namespace lxb {

constexpr int kTriWarps = 4;  // warps per CTA
constexpr int kTile = 32;

template <typename T>
__global__ void __launch_bounds__(kTriWarps * 32)
    tridiagonal_kernel(const T* __restrict__ D, const T* __restrict__ DL, const T* __restrict__ DU,
                       int64_t sD, const T* __restrict__ B, int64_t sB, T* __restrict__ X,
                       T* __restrict__ ws, int64_t batch, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* tiles = reinterpret_cast<T*>(smem_raw) + (size_t)warp * 4 * kTile * 33;
  T* td = tiles;
  T* tl = td + kTile * 33;
  T* tu = tl + kTile * 33;
  T* tb = tu + kTile * 33;
  const int64_t wslot = (int64_t)blockIdx.x * kTriWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kTriWarps;
  T* w1 = ws + wslot * 2 * (int64_t)n * 32;  // u1[i][lane]
  T* w2 = w1 + (int64_t)n * 32;              // u2[i][lane]
  const int64_t sOff = sD ? sD - 1 : 0;      // off-diagonals hold n-1 entries per system
  const int64_t groups = (batch + 31) / 32;
  const int nchunks = (n + kTile - 1) / kTile;

  for (int64_t g = wslot; g < groups; g += nwarps) {
    const int64_t sys0 = g * 32;
    const int nsys = (int)((batch - sys0) < 32 ? (batch - sys0) : 32);
    // carried state of the current row (row i), initialised at the first chunk
    T cd = T(0), cu = T(0), cb = T(0);
    // -------- forward sweep
    for (int c = 0; c < nchunks; ++c) {
      const int i0 = c * kTile;
      __syncwarp();
#pragma unroll 4
      for (int r = 0; r < nsys; ++r) {
        const int64_t s = sys0 + r;
        const int i = i0 + lane;
        const T vd = i < n ? D[s * sD + i] : T(1);
        const T vl = (i >= 1 && i < n) ? DL[s * sOff + i - 1] : T(0);  // sub-diagonal entry of row i
        const T vu = i < n - 1 ? DU[s * sOff + i] : T(0);
        const T vb = i < n ? B[s * sB + i] : T(0);
        td[lane * 33 + r] = vd;
        tl[lane * 33 + r] = vl;
        tu[lane * 33 + r] = vu;
        tb[lane * 33 + r] = vb;
      }
      __syncwarp();
      if (lane < nsys) {
        const int cnt = (n - i0) < kTile ? (n - i0) : kTile;
        for (int e = 0; e < cnt; ++e) {
          const int i = i0 + e;
          if (i == 0) {
            cd = td[lane];
            cu = tu[lane];
            cb = tb[lane];
            continue;
          }
          // eliminate the sub-diagonal entry of row i against the carried row i-1
          const T dl = tl[e * 33 + lane], dn = td[e * 33 + lane], un = tu[e * 33 + lane],
                  bn = tb[e * 33 + lane];
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
          // y_{i-1} goes to the output tile slot of element e-1 (previous chunk's last element is
          // written through td, see below)
          if (e > 0) tb[(e - 1) * 33 + lane] = ob * rinv;
          else td[lane] = ob * rinv;  // belongs to element i0-1 of the previous chunk
        }
        if (i0 + cnt == n) tb[(cnt - 1) * 33 + lane] = cb / cd;  // last row: y_{n-1} = x_{n-1}
      }
      __syncwarp();
      // coalesced store of y for this chunk (elements i0 .. i0+cnt-1, except the chunk's last
      // element which is only known after the next chunk's first step) and of element i0-1
      for (int r = 0; r < nsys; ++r) {
        const int64_t s = sys0 + r;
        const int i = i0 + lane;
        const bool last_chunk = i0 + kTile >= n;
        const int upto = last_chunk ? n : i0 + kTile - 1;  // exclusive bound of final values
        if (i < upto) X[s * n + i] = tb[lane * 33 + r];
        if (lane == 0 && i0 > 0) X[s * n + i0 - 1] = td[r];
      }
    }
    // -------- backward sweep: x_i = y_i - u1_i x_{i+1} - u2_i x_{i+2}
    T x1 = T(0), x2 = T(0);  // x_{i+1}, x_{i+2}
    for (int c = nchunks - 1; c >= 0; --c) {
      const int i0 = c * kTile;
      __syncwarp();
      for (int r = 0; r < nsys; ++r) {
        const int i = i0 + lane;
        tb[lane * 33 + r] = i < n ? X[(sys0 + r) * n + i] : T(0);
      }
      __syncwarp();
      if (lane < nsys) {
        const int cnt = (n - i0) < kTile ? (n - i0) : kTile;
        for (int e = cnt - 1; e >= 0; --e) {
          const int i = i0 + e;
          T xi = tb[e * 33 + lane];
          if (i < n - 1) {
            const T u1 = w1[(int64_t)i * 32 + lane], u2 = w2[(int64_t)i * 32 + lane];
            xi = xi - u1 * x1 - u2 * x2;
          }
          tb[e * 33 + lane] = xi;
          x2 = x1;
          x1 = xi;
        }
      }
      __syncwarp();
      for (int r = 0; r < nsys; ++r) {
        const int i = i0 + lane;
        if (i < n) X[(sys0 + r) * n + i] = tb[lane * 33 + r];
      }
    }
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
  pl.smem = (size_t)kTriWarps * 4 * kTile * 33 * sizeof(T);
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

// LU with partial pivoting for vmapped 32x32 fp32 systems (BASELINE configs[1]):
// lineax/_solver/lu.py:43-66 (jsp.linalg.lu_factor / lu_solve), arithmetic order of oracle/getf2.c.
//
// One warp per system, lane r owns row r in 16 packed f32x2 registers.
//  * A is staged HBM -> shared memory by TMA (`cp.async.bulk.tensor.3d`, SWIZZLE_128B, one 4 KB box
//    per system, completion on a per-warp mbarrier); the next system of the warp is in flight while
//    the current one is eliminated.  The 128-byte swizzle makes the row-per-lane LDS.128 reads
//    conflict-free.
//  * Step k: every lane turns |a_rk| into an integer key (candidates carry the top bit, every NaN
//    maps to one key so that, like ISAMAX, the first NaN wins), one `redux.sync.max` finds the
//    maximum, a second `redux.sync.min` over the physical row positions of the lanes holding it
//    applies LAPACK's first-index tie-break and at the same time yields piv[k].  Each lane forms
//    1/|a_rk| while the reductions are in flight (IEEE division, sign restored by the pivot lane).
//  * The pivot lane writes its row (columns >= k) into line k of a per-warp 4 KB "U buffer" and
//    {1/pivot, y_k} into slot k of an extras array; everybody reads the line back with broadcast
//    LDS.128 and does the rank-1 update with `fma.rn.f32x2` (FFMA2; IEEE per element, so results
//    stay bit-identical to the scalar oracle).  Because line k is never rewritten, after the last
//    step the U buffer holds U in LAPACK's final row order and the extras hold 1/u_kk and the
//    forward-substituted right-hand side: the back substitution needs no permutation, lane r
//    reloads row r of U and the x_k are passed around with one shuffle per step.
//  * Optional (lu, piv) output: every lane drops its full row into the U buffer at its final
//    position and one TMA store writes the 4 KB tile.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "lu_tma.cuh"

namespace lxb {
namespace {

constexpr int kTile = 32 * 32 * 4;  // bytes of one system
constexpr int kExt = 32 * 8;        // {1/pivot, y_k} per elimination step

// U buffer layouts.  Line k only ever carries the 16-byte chunks with a column > k, so when the (lu, piv)
// state is not written the lines are packed back to back (2176 B instead of 4096 B per warp, which is
// what lets a fourth CTA fit on an SM); with state output the buffer is the full swizzled tile that the
// TMA store writes out.
__host__ __device__ constexpr int tri_first_chunk(int k) { return (k + 1) / 4; }
__host__ __device__ constexpr int tri_line_off(int k) {  // in 16-byte chunks
  int off = 0;
  for (int j = 0; j < k; ++j) off += 8 - tri_first_chunk(j);
  return off;
}
constexpr int kTriBytes = tri_line_off(32) * 16;  // 2176
template <bool FULL>
__host__ __device__ constexpr int ubuf_bytes() { return FULL ? kTile : kTriBytes; }
// per warp: U buffer + extras + mbarrier (padded so that a full-tile U buffer stays 1 KB aligned)
template <bool FULL>
__host__ __device__ constexpr int warp_small_bytes() { return FULL ? kTile + 1024 : kTriBytes + kExt + 16; }
// byte offset of chunk c (>= first chunk) of line k
template <bool FULL>
__host__ __device__ constexpr int ubuf_off(int k, int c) {
  return FULL ? k * 128 + ((c ^ (k & 7)) << 4) : (tri_line_off(k) + c - tri_first_chunk(k)) * 16;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!__all_sync(kFull, ok));  // warp-uniform exit: keeps the warp provably converged
}
__device__ __forceinline__ void tma_load_sys(uint32_t dst, const CUtensorMap* tm, int sys,
                                             uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tm), "r"(0), "r"(0), "r"(sys), "r"(mbar)
      : "memory");
}
__device__ __forceinline__ void tma_store_sys(const CUtensorMap* tm, int sys, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm),
               "r"(0), "r"(0), "r"(sys), "r"(src)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// d = a * b + c on both halves, round-to-nearest-even per element (FFMA2)
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void lds128(uint32_t addr, uint64_t& x, uint64_t& y) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint64_t x, uint64_t y) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(x), "l"(y) : "memory");
}
__device__ __forceinline__ float elem(const uint64_t (&a)[16], int j) {
  float lo, hi;
  unpack2(a[j >> 1], lo, hi);
  return (j & 1) ? hi : lo;
}

// cold path of the reciprocal (zero / denormal / huge / inf / NaN pivot): full IEEE division
__device__ __noinline__ float rcp_slow(float x) { return 1.0f / x; }

struct LuLane {
  int pos;    // physical (LAPACK) position of the row this lane holds, while it is a pivot candidate
  int fpos;   // final position (WLU only)
  int mypiv;  // piv[lane] (WLU only)
  bool done;  // row already chosen as a pivot
  float bb;   // right-hand side entry of the row (forward substitution fused in the elimination)
};

// One elimination step (column K).
template <int K, bool SOLVE, bool WLU>
__device__ __forceinline__ void lu_step(uint64_t (&a)[16], LuLane& s, uint32_t ubuf, uint32_t ext, int lane) {
  const float v = elem(a, K);
  unsigned ab = __float_as_uint(v) & 0x7fffffffu;
  ab = min(ab, 0x7f800001u);  // every NaN -> one key: like ISAMAX the first NaN wins
  const int key = s.done ? -1 : (int)ab;
  // every lane forms 1 / |candidate| while the reductions are in flight (same issue slots as doing it
  // in the pivot lane only, but off the critical path).  MUFU.RCP + one Newton step is the correctly
  // rounded reciprocal for |x| in [2^-126, 2^126); the pivot lane re-does the rare rest with a full
  // IEEE division.
  const float x = __uint_as_float(ab);
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(x));
  float rabs = fma_(r0, fma_(-x, r0, 1.0f), r0);
  asm volatile("" : "+f"(rabs));  // keep it here: do not sink it into the pivot-lane branch
  const int m = __reduce_max_sync(kFull, key);
  const int pc = key == m ? s.pos : 64;
  const int pm = __reduce_min_sync(kFull, pc);  // first-index tie-break; = piv[K]
  const bool is = pc == pm;
  constexpr int C0 = tri_first_chunk(K);  // first 16-byte chunk with a column > K
  if (is) {
    // (tested on the warp-uniform maximum m = the pivot's |value| bits, so the branch is uniform)
    if ((unsigned)m - 0x00800000u >= 0x7e000000u) rabs = rcp_slow(x);
    const float r = __uint_as_float(__float_as_uint(rabs) | (__float_as_uint(v) & 0x80000000u));
#pragma unroll
    for (int c = C0; c < 8; ++c) sts128(ubuf + ubuf_off<WLU>(K, c), a[2 * c], a[2 * c + 1]);
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ext + K * 8), "f"(r), "f"(s.bb) : "memory");
  }
  __syncwarp();
  uint64_t u[16];
#pragma unroll
  for (int c = C0; c < 8; ++c) lds128(ubuf + ubuf_off<WLU>(K, c), u[2 * c], u[2 * c + 1]);
  float r, bpk;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r), "=f"(bpk) : "r"(ext + K * 8));
  if (WLU) {
    if (lane == K) s.mypiv = pm;
    s.fpos = is ? K : s.fpos;
  }
  s.pos = s.pos == K ? pm : s.pos;
  s.done = s.done || is;
  if (!s.done) {
    // rank-1 update of this row (and of its right-hand side entry): a_rj = fma(-l, u_kj, a_rj)
    const float nl = v * (-r);
    float lo, hi;
    unpack2(a[K >> 1], lo, hi);
    if ((K & 1) == 0) {
      float ulo, uhi;
      unpack2(u[K >> 1], ulo, uhi);
      hi = fma_(nl, uhi, hi);
      if (WLU) lo = -nl;
    } else {
      if (WLU) hi = -nl;
    }
    a[K >> 1] = pack2(lo, hi);
    const uint64_t nl2 = pack2(nl, nl);
#pragma unroll
    for (int p = (K >> 1) + 1; p < 16; ++p) a[p] = ffma2(nl2, u[p], a[p]);
    if (SOLVE) s.bb = fma_(nl, bpk, s.bb);
  }
}

template <int K, bool SOLVE, bool WLU>
__device__ __forceinline__ void lu_steps(uint64_t (&a)[16], LuLane& s, uint32_t ubuf, uint32_t ext, int lane) {
  if constexpr (K < 32) {
    lu_step<K, SOLVE, WLU>(a, s, ubuf, ext, lane);
    lu_steps<K + 1, SOLVE, WLU>(a, s, ubuf, ext, lane);
  }
}

// UNI: every warp of a CTA runs the same number of iterations (a warp past the end of the batch redoes
// the last system with its stores masked), which lets the compiler prove warp convergence at the
// redux / shuffle collectives and drop the divergence wrappers around them.
template <int WARPS, bool SOLVE, bool WLU, bool UNI, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    lu32_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmLU,
                    const float* __restrict__ B, int64_t sB, float* __restrict__ X,
                    int32_t* __restrict__ PIV, int64_t batch) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // [WARPS x 4 KB TMA stage (SWIZZLE_128B boxes: 1 KB aligned)] [WARPS x {U buffer, extras, mbarrier}]:
  // the hot loop addresses everything of a warp from ONE register (ubuf) with constant offsets
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t stage = base + warp * kTile;
  constexpr int UB = ubuf_bytes<WLU>();
  uint32_t ubuf = base + WARPS * kTile + warp * warp_small_bytes<WLU>();
  asm volatile("mov.u32 %0, %0;" : "+r"(ubuf));  // opaque: keep it in a register, never rematerialise
  const uint32_t ext = ubuf + UB;
  const uint32_t mbar = ext + kExt;
  if (lane == 0) mbar_init(mbar, 1);
  __syncwarp();
  const int64_t nblk = (batch + WARPS - 1) / WARPS;
  {
    const int64_t first = (int64_t)blockIdx.x * WARPS + warp;
    if (lane == 0 && blockIdx.x < nblk && (UNI || first < batch)) {
      mbar_expect_tx(mbar, kTile);
      tma_load_sys(stage, &tmA, (int)(first < batch ? first : batch - 1), mbar);
    }
  }
  uint32_t parity = 0;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const bool valid = blk * WARPS + warp < batch;
    if (!UNI && !valid) break;
    const int64_t sys = valid ? blk * WARPS + warp : batch - 1;
    float bb = 0.f;
    if (SOLVE) bb = B[sys * sB + lane];
    mbar_wait(mbar, parity);
    parity ^= 1u;
    uint64_t a[16];
    {
      const uint32_t row = stage + lane * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) lds128(row + ((c ^ (lane & 7)) << 4), a[2 * c], a[2 * c + 1]);
    }
    __syncwarp();
    if (lane == 0) {
      const int64_t nb = blk + gridDim.x;
      const int64_t nxt = nb * WARPS + warp;
      if (nb < nblk && (UNI || nxt < batch)) {
        fence_proxy_async();  // the generic-proxy reads above precede the async-proxy refill
        mbar_expect_tx(mbar, kTile);
        tma_load_sys(stage, &tmA, (int)(nxt < batch ? nxt : batch - 1), mbar);
      }
      if (WLU) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // U buffer free again
    }
    if (WLU) __syncwarp();

    // ---- elimination ----
    LuLane s;
    s.pos = lane;
    s.fpos = lane;
    s.mypiv = lane;
    s.done = false;
    s.bb = bb;
    lu_steps<0, SOLVE, WLU>(a, s, ubuf, ext, lane);

    if (WLU) {
      // every lane drops its complete row (L | U) at its final position, one TMA store writes the tile
      __syncwarp();
      const uint32_t row = ubuf + s.fpos * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) sts128(row + ((c ^ (s.fpos & 7)) << 4), a[2 * c], a[2 * c + 1]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && valid) tma_store_sys(&tmLU, (int)sys, ubuf);
      if (PIV != nullptr && valid) PIV[sys * 32 + lane] = s.mypiv;
    }
    if (SOLVE) {
      // ---- back substitution on U (final row order) ----
      if constexpr (WLU) {
        const uint32_t row = ubuf + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) lds128(row + ((c ^ (lane & 7)) << 4), a[2 * c], a[2 * c + 1]);
      } else {
        // packed lines: row `lane` starts at chunk 8*lane - sum_{i<=lane} floor(i/4) and holds the
        // chunks from (lane+1)/4 on -- exactly the columns > lane the substitution reads
        const int q = lane >> 2, r = lane & 3, c0 = (lane + 1) >> 2;
        const uint32_t row = ubuf + (8 * lane - (2 * q * (q - 1) + q * (r + 1)) - c0) * 16;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c >= c0) lds128(row + c * 16, a[2 * c], a[2 * c + 1]);
      }
      float rinv, y;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(rinv), "=f"(y) : "r"(ext + lane * 8));
#pragma unroll
      for (int k = 31; k >= 1; --k) {
        const float xk = __shfl_sync(kFull, y * rinv, k);
        if (lane < k) y = fma_(-elem(a, k), xk, y);
      }
      if (valid) X[sys * 32 + lane] = y * rinv;
    }
    __syncwarp();  // U buffer / extras are rewritten by the next system
  }
  if (WLU && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// [batch, 32, 32] f32 with an element stride between systems, one swizzled 32x32 box per system
int make_sys_map(CUtensorMap* tm, const float* ptr, int64_t stride, int64_t batch) {
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return LXB_E_UNSUPPORTED;
  const cuuint64_t dims[3] = {32, 32, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {128, (cuuint64_t)stride * 4};
  const cuuint32_t box[3] = {32, 32, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : LXB_E_UNSUPPORTED;
}

template <int WARPS, bool SOLVE, bool WLU, bool UNI, int MINB>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmLU, const float* b, int64_t sb, float* x,
           int32_t* piv, int64_t batch, cudaStream_t st) {
  auto kern = lu32_tma_kernel<WARPS, SOLVE, WLU, UNI, MINB>;
  const size_t smem = (size_t)WARPS * (kTile + warp_small_bytes<WLU>());
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) occ = 1;
  int64_t blocks = (batch + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)kNumSMs * occ;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(tmA, tmLU, b, sb, x, piv, batch);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool lu32_tma_eligible(const float* A, int64_t sA, const float* lu, int64_t batch, int n) {
  if (n != 32 || batch < 1 || batch > 0x7fffffff) return false;
  if (!aligned16(A) || sA < 1024 || (sA % 4) != 0) return false;
  if (lu != nullptr && !aligned16(lu)) return false;
  return encode_tiled() != nullptr;
}

int lu32_tma_launch(const float* A, int64_t sA, const float* b, int64_t sb, float* x, float* lu,
                    int32_t* piv, int64_t batch, bool solve, cudaStream_t st) {
  CUtensorMap tmA, tmLU;
  int rc = make_sys_map(&tmA, A, sA, batch);
  if (rc != 0) return rc;
  if (lu != nullptr) {
    rc = make_sys_map(&tmLU, lu, 1024, batch);
    if (rc != 0) return rc;
  } else {
    tmLU = tmA;
  }
  constexpr int W = 8;
  // tuning knobs (see lu32_tma_kernel): LXB_LU_NONUNI=1, LXB_LU_MINB=3 (80 registers, 3 CTAs per SM)
  static const bool uni = getenv("LXB_LU_NONUNI") == nullptr;
  static const bool minb3 = getenv("LXB_LU_MINB") != nullptr && atoi(getenv("LXB_LU_MINB")) == 3;
  if (solve) {
    if (lu != nullptr)
      return uni ? launch<W, true, true, true, 3>(tmA, tmLU, b, sb, x, piv, batch, st)
                 : launch<W, true, true, false, 3>(tmA, tmLU, b, sb, x, piv, batch, st);
    if (minb3)
      return uni ? launch<W, true, false, true, 3>(tmA, tmLU, b, sb, x, piv, batch, st)
                 : launch<W, true, false, false, 3>(tmA, tmLU, b, sb, x, piv, batch, st);
    return uni ? launch<W, true, false, true, 4>(tmA, tmLU, b, sb, x, piv, batch, st)
               : launch<W, true, false, false, 4>(tmA, tmLU, b, sb, x, piv, batch, st);
  }
  return uni ? launch<W, false, true, true, 3>(tmA, tmLU, b, sb, x, piv, batch, st)
             : launch<W, false, true, false, 3>(tmA, tmLU, b, sb, x, piv, batch, st);
}

}  // namespace lxb

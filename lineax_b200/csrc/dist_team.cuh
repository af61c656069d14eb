// Cross-GPU team for the row-sharded persistent solver kernels (gmres_dist.cu, lsmr_dist.cu):
// a GridTeam per GPU plus exchanges over NVLink peer memory (symmetric buffers).  No NCCL call and
// no host round trip on the data path.
//
// One exchange round (`xround`) costs one NVLink one-way latency plus a local fold, with NO grid barrier:
//   every CTA drops its partials into a local buffer and bumps an arrival counter; the LAST CTA to
//   arrive folds the partials (fixed order => deterministic) and writes this GPU's totals straight into
//   a slot of EVERY peer as self-validating words -- each 8-byte word carries 32 bits of payload and the
//   32-bit round number (the "LL" scheme of collective libraries), so no separate flag, fence or second
//   hop is needed; every CTA of every GPU then simply reads the K x P words of its own buffer until
//   their round number matches and folds the P contributions in rank order (identical bits everywhere).
// Round 1 measured ~25 us per round with a cooperative grid barrier + CTA-0 fold + flag round trip.
#pragma once
#include "krylov_grid.cuh"
#include "krylov_grid_api.cuh"

namespace lxb {

constexpr int kMaxPeers = 16;
constexpr int kDistThreads = 512;  // 16 warps per CTA: one CTA per SM when x is staged in shared memory
constexpr size_t kSymmFlagBytes = 4096;  // [0]: persistent epoch, [64 + 16*r]: arrival slot of rank r

__host__ __device__ inline size_t symm_part_off() { return kSymmFlagBytes; }
// two rounds in flight x (kGridMaxK + 1 barrier word) values x kMaxPeers ranks x 8-byte words
// (a double travels as two words)
template <typename T>
__host__ __device__ inline size_t symm_xchg_off() {
  return kSymmFlagBytes +
         (((size_t)2 * (kGridMaxKHost + 1) * kMaxPeers * 2 * sizeof(T)) + 255) / 256 * 256;
}
template <typename T>
size_t symm_bytes(int n) { return symm_xchg_off<T>() + pad4(n) * sizeof(T) + 256; }

template <typename T>
struct DistTeam {
  GridTeam<T> g;
  int P, rank;
  unsigned char* const* peers;  // device array of P symmetric-buffer base pointers
  unsigned char* mine;
  unsigned long long epoch;
  int xflip;

  __device__ DistTeam(T* part, T* red, unsigned char* const* peers_, int P_, int rank_)
      : g(part, red), P(P_), rank(rank_), peers(peers_), mine(peers_[rank_]), xflip(0) {
    epoch = *reinterpret_cast<volatile unsigned long long*>(mine);
  }

  static constexpr int kWords = sizeof(T) / 4;  // 8-byte {payload, round} words per value

  __device__ __forceinline__ static void ll_store(unsigned long long* dst, T v, unsigned tag) {
    if constexpr (sizeof(T) == 4) {
      const unsigned long long w = ((unsigned long long)tag << 32) | __float_as_uint(v);
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w) : "memory");
    } else {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
      const unsigned long long w0 = ((unsigned long long)tag << 32) | (bits & 0xffffffffull);
      const unsigned long long w1 = ((unsigned long long)tag << 32) | (bits >> 32);
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w0) : "memory");
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(w1) : "memory");
    }
  }
  __device__ __forceinline__ static T ll_wait(const unsigned long long* src, unsigned tag) {
    unsigned long long w0, w1 = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w0) : "l"(src) : "memory");
    } while ((unsigned)(w0 >> 32) != tag);
    if constexpr (sizeof(T) == 4) {
      return __uint_as_float((unsigned)w0);
    } else {
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w1) : "l"(src + 1) : "memory");
      } while ((unsigned)(w1 >> 32) != tag);
      return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    }
  }

  // ONE fused round = grid-wide + cross-GPU all-reduce of KS sums and KM abs-maxima AND a cross-GPU
  // barrier for every remote store issued before it (vector pushes); see the header comment.
  // `sums` / `maxes` live in shared memory and hold this CTA's block-reduced values on entry,
  // the global results on exit (identical bits on every CTA of every GPU).
  __device__ void xround(T* sums, int KS, T* maxes, int KM) {
    const int K = KS + KM;  // + 1 barrier word, so that a round with K = 0 still synchronises
    T* lbuf = g.part + (size_t)g.flip * kGridMaxK * g.nb;
    g.flip ^= 1;
    const int par = xflip;
    const size_t off = symm_part_off() + (size_t)par * (kGridMaxK + 1) * kMaxPeers * 8 * kWords;
    xflip ^= 1;
    epoch += 1;
    const unsigned tag = (unsigned)epoch;
    unsigned int* cnt = reinterpret_cast<unsigned int*>(mine + 1024) + par;  // arrival counter of this parity
    __shared__ int s_last;
    __shared__ T xstage[kGridMaxK * kMaxPeers];  // the round's P contributions per value
    auto stage_buf = [&](int k, int q) -> T& { return xstage[k * kMaxPeers + q]; };
    __syncthreads();
    for (int k = g.tid; k < K; k += g.nt) lbuf[(size_t)k * g.nb + g.bid] = k < KS ? sums[k] : maxes[k - KS];
    __threadfence_system();  // partials visible device-wide, earlier remote pushes system-wide
    __syncthreads();
    if (g.tid == 0) {
      const unsigned old = atomicAdd(cnt, 1u);
      s_last = old == (unsigned)g.nb - 1u;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();  // acquire: every CTA's partials (and its pushes) happen-before what follows
      const int lane = g.tid & 31, warp = g.tid >> 5, nw = g.nt >> 5;
      for (int k = warp; k <= K; k += nw) {
        T a = T(0);
        if (k < KS) {
          for (int i = lane; i < g.nb; i += 32) a += __ldcg(lbuf + (size_t)k * g.nb + i);
          a = warp_sum(a);
        } else if (k < K) {
          for (int i = lane; i < g.nb; i += 32) a = absmax2(a, __ldcg(lbuf + (size_t)k * g.nb + i));
          a = warp_absmax(a);
        }
        if (lane < P)
          ll_store(reinterpret_cast<unsigned long long*>(peers[lane] + off) + ((size_t)k * kMaxPeers + rank) * kWords,
                   a, tag);
      }
      if (g.tid == 0) *cnt = 0u;  // everybody of this round has arrived; the parity is reused two rounds on
    }
    // every CTA: wait for the (K + 1) x P self-validating words of this round, fold in rank order
    const unsigned long long* in = reinterpret_cast<const unsigned long long*>(mine + off);
    for (int idx = g.tid; idx < (K + 1) * P; idx += g.nt) {
      const int k = idx / P, q = idx % P;
      const T v = ll_wait(in + ((size_t)k * kMaxPeers + q) * kWords, tag);
      if (k < K) stage_buf(k, q) = v;
    }
    __syncthreads();
    for (int k = g.tid; k < K; k += g.nt) {
      T a = T(0);
      if (k < KS) {
        for (int q = 0; q < P; ++q) a += stage_buf(k, q);
        sums[k] = a;
      } else {
        for (int q = 0; q < P; ++q) a = absmax2(a, stage_buf(k, q));
        maxes[k - KS] = a;
      }
    }
    __syncthreads();
  }

  // push my slice [lo, hi) (local indices) of a vector into every GPU's exchange buffer at the
  // global position; readable locally at xchg() after the next xround().
  __device__ T* xchg() const { return reinterpret_cast<T*>(mine + symm_xchg_off<T>()); }
  __device__ void push(const T* local, int lo, int hi, int row_offset) {
    for (int q = 0; q < P; ++q) {
      T* dst = reinterpret_cast<T*>(peers[q] + symm_xchg_off<T>()) + row_offset;
      for (int i = lo + g.tid; i < hi; i += g.nt) dst[i] = local[i];
    }
  }
  __device__ void finish(T* scratch) {
    xround(scratch, 0, scratch, 0);  // nobody leaves (and reuses exchange slots) while a peer still reads
    g.sync();
    if (g.bid == 0 && g.tid == 0) *reinterpret_cast<volatile unsigned long long*>(mine) = epoch;
  }
};

}  // namespace lxb

// Cross-GPU team for the row-sharded persistent solver kernels (gmres_dist.cu, lsmr_dist.cu):
// a GridTeam per GPU plus exchanges over NVLink peer memory (symmetric buffers).  No NCCL call and
// no host round trip on the data path.  The cross-GPU barrier is a monotonically increasing epoch
// written with system-scope release stores into every peer's flag array and polled locally.
#pragma once
#include "krylov_grid.cuh"
#include "krylov_grid_api.cuh"

namespace lxb {

constexpr int kMaxPeers = 16;
constexpr int kDistThreads = 512;  // 16 warps per CTA: one CTA per SM when x is staged in shared memory
constexpr size_t kSymmFlagBytes = 4096;  // [0]: persistent epoch, [64 + 16*r]: arrival slot of rank r

__host__ __device__ inline size_t symm_part_off() { return kSymmFlagBytes; }
template <typename T>
__host__ __device__ inline size_t symm_xchg_off() {
  return kSymmFlagBytes + (((size_t)2 * kGridMaxKHost * kMaxPeers * sizeof(T)) + 255) / 256 * 256;
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

  // ONE fused round = grid-wide + cross-GPU all-reduce of KS sums and KM abs-maxima AND a
  // cross-GPU barrier for every remote store issued before it (vector pushes):
  //   per-CTA partials -> local buffer -> grid barrier -> CTA 0 folds them and pushes this GPU's
  //   totals into a slot of every peer -> epoch flags (release.sys) -> every CTA polls its own
  //   GPU's flags (acquire.sys) -> every CTA folds the P slots in rank order.
  // `sums` / `maxes` live in shared memory and hold this CTA's block-reduced values on entry,
  // the global results on exit (identical bits on every CTA of every GPU).
  __device__ void xround(T* sums, int KS, T* maxes, int KM) {
    const int K = KS + KM;
    T* lbuf = g.part + (size_t)g.flip * kGridMaxK * g.nb;
    g.flip ^= 1;
    const size_t off = symm_part_off() + (size_t)xflip * kGridMaxK * kMaxPeers * sizeof(T);
    xflip ^= 1;
    __syncthreads();
    for (int k = g.tid; k < K; k += g.nt) lbuf[(size_t)k * g.nb + g.bid] = k < KS ? sums[k] : maxes[k - KS];
    __threadfence_system();  // also orders this thread's earlier remote pushes
    g.sync();
    epoch += 1;
    if (g.bid == 0) {
      const int lane = g.tid & 31, warp = g.tid >> 5, nw = g.nt >> 5;
      for (int k = warp; k < K; k += nw) {
        T a = T(0);
        if (k < KS) {
          for (int i = lane; i < g.nb; i += 32) a += __ldcg(lbuf + (size_t)k * g.nb + i);
          a = warp_sum(a);
        } else {
          for (int i = lane; i < g.nb; i += 32) a = absmax2(a, __ldcg(lbuf + (size_t)k * g.nb + i));
          a = warp_absmax(a);
        }
        if (lane < P) reinterpret_cast<T*>(peers[lane] + off)[(size_t)k * kMaxPeers + rank] = a;
      }
      __threadfence_system();
      __syncthreads();
      if (g.tid < P) {
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(peers[g.tid] + 64 + 16 * rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(epoch) : "memory");
      }
    }
    if (g.tid < P) {
      const unsigned long long* my = reinterpret_cast<const unsigned long long*>(mine + 64 + 16 * g.tid);
      unsigned long long seen;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(my) : "memory");
      } while (seen < epoch);
    }
    __syncthreads();
    const T* in = reinterpret_cast<const T*>(mine + off);
    for (int k = g.tid; k < K; k += g.nt) {
      T a = T(0);
      if (k < KS) {
        for (int q = 0; q < P; ++q) a += __ldcg(in + (size_t)k * kMaxPeers + q);
        sums[k] = a;
      } else {
        for (int q = 0; q < P; ++q) a = absmax2(a, __ldcg(in + (size_t)k * kMaxPeers + q));
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

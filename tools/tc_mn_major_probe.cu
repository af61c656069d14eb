// Standalone probe for the MN-major 3xTF32 tcgen05 GEMM used by the (not yet correct) tensor-core
// W = V^T A2 kernel of profiles/r01_wtc_attempt.patch.  One CTA computes W[32][128] = V^T A2 for
// V (R x 32) and A2 (R x 128) with R = 32 * nchunks rows and the result is compared with a double
// precision host product, for several chunk counts.  Build and run on a B200:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tc_mn_major_probe.cu -o /tmp/probe && /tmp/probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int sw128(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ void split_store(unsigned char* hi, unsigned char* lo, int r, int c, float4 x) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = x.x - h.x;
  h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = x.y - h.y;
  h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = x.z - h.z;
  h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = x.w - h.w;
  *reinterpret_cast<float4*>(hi + sw128(r, c)) = h;
  *reinterpret_cast<float4*>(lo + sw128(r, c)) = l;
}
constexpr uint32_t kIdescW = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) |
                             ((128u >> 4) << 24);
// experiment knobs (set per launch): instruction descriptor, LBO / SBO bytes, layout type
__device__ uint32_t g_idesc = kIdescW, g_lbo = 4096, g_sbo = 1024, g_layout = 2;
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((g_lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((g_sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)g_layout << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(da), "l"(db), "r"(g_idesc), "r"(acc) : "memory");
}

// A2: [R][lda] (128 columns used), V: [R][ldv] (32 columns used), W: [32][128]
__global__ void __launch_bounds__(128) probe_kernel(const float* A2, int lda, const float* V, int ldv, float* W, int R,
                                                    int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *Ahi = base, *Alo = base + 16384, *Bhi = base + 32768, *Blo = base + 36864;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(base + 40960);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + 40976);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
  float acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.f;
  uint32_t phase = 0;
  const int nchunks = R / 32;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int rb = ch * 32;
    for (int i = 0; i < 8; ++i) {
      const int p = tid + i * 128, r = p >> 5, c16 = p & 31, g = c16 >> 3, c = c16 & 7;
      split_store(Ahi + g * 4096, Alo + g * 4096, r, c, reinterpret_cast<const float4*>(A2 + (size_t)(rb + r) * lda)[c16]);
    }
    for (int i = 0; i < 2; ++i) {
      const int p = tid + i * 128, r = p >> 3, c = p & 7;
      split_store(Bhi, Blo, r, c, reinterpret_cast<const float4*>(V + (size_t)(rb + r) * ldv)[c]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t dah = desc_mn(a_hi + 1024 * kb, 4096), dal = desc_mn(a_lo + 1024 * kb, 4096);
        const uint64_t dbh = desc_mn(b_hi + 1024 * kb, 4096), dbl = desc_mn(b_lo + 1024 * kb, 4096);
        mma(tmem, dal, dbh, kb > 0 ? 1u : 0u);
        mma(tmem, dah, dbl, 1u);
        mma(tmem, dah, dbh, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
    if (!done && tid == 0) atomicExch(err, 1);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
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
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] += __uint_as_float(r[k]);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) W[k * 128 + tid] = acc[k];
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

// Control: the K-major configuration of qr_update_tc_kernel (known good): P[128][128] = A[128][32] * B[128][32]^T
constexpr uint32_t kIdescK = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__global__ void __launch_bounds__(128) control_kernel(const float* A, const float* B, float* P, int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *Bhi = base, *Blo = base + 16384, *Ahi = base + 32768, *Alo = base + 49152;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(base + 65536);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + 65536 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int c = 0; c < 8; ++c) {
    split_store(Ahi, Alo, tid, c, reinterpret_cast<const float4*>(A + tid * 32)[c]);
    split_store(Bhi, Blo, tid, c, reinterpret_cast<const float4*>(B + tid * 32)[c]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t dah = desc_k(smem_u32(Ahi) + 32 * ks), dal = desc_k(smem_u32(Alo) + 32 * ks);
      const uint64_t dbh = desc_k(smem_u32(Bhi) + 32 * ks), dbl = desc_k(smem_u32(Blo) + 32 * ks);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(dal), "l"(dbh), "r"(kIdescK), "r"(ks > 0 ? 1u : 0u) : "memory");
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(dah), "l"(dbl), "r"(kIdescK), "r"(1u) : "memory");
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(dah), "l"(dbh), "r"(kIdescK), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
  }
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(0u) : "memory");
  if (!done && tid == 0) atomicExch(err, 1);
  asm volatile("tcgen05.fence::after_thread_sync;");
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
#pragma unroll
    for (int k = 0; k < 32; ++k) P[tid * 128 + 32 * qc + k] = __uint_as_float(r[k]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// Candidate fix: keep BOTH operands K-major (the verified configuration) by transposing the chunks while staging.
// Thread (rq = tid & 7, cq) loads a 4 x 4 block (rows 4rq.., columns 4cq..), transposes it in registers and stores
// four 16-byte K-chunks (one per column); a quarter warp covers the 8 chunks of one tile row: conflict free.
constexpr uint32_t kIdescKT = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__global__ void __launch_bounds__(128) probe_kernel_kt(const float* A2, int lda, const float* V, int ldv, float* W, int R,
                                                       int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *Ahi = base, *Alo = base + 16384, *Bhi = base + 32768, *Blo = base + 36864;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(base + 40960);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + 40976);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  float acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.f;
  uint32_t phase = 0;
  const int rq = tid & 7;
  for (int ch = 0; ch < R / 32; ++ch) {
    const int rb = ch * 32;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int cq = (tid >> 3) + 16 * i;
      float4 x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = reinterpret_cast<const float4*>(A2 + (size_t)(rb + 4 * rq + j) * lda)[cq];
      split_store(Ahi, Alo, 4 * cq + 0, rq, make_float4(x[0].x, x[1].x, x[2].x, x[3].x));
      split_store(Ahi, Alo, 4 * cq + 1, rq, make_float4(x[0].y, x[1].y, x[2].y, x[3].y));
      split_store(Ahi, Alo, 4 * cq + 2, rq, make_float4(x[0].z, x[1].z, x[2].z, x[3].z));
      split_store(Ahi, Alo, 4 * cq + 3, rq, make_float4(x[0].w, x[1].w, x[2].w, x[3].w));
    }
    if (tid < 64) {
      const int cq = tid >> 3;
      float4 x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = reinterpret_cast<const float4*>(V + (size_t)(rb + 4 * rq + j) * ldv)[cq];
      split_store(Bhi, Blo, 4 * cq + 0, rq, make_float4(x[0].x, x[1].x, x[2].x, x[3].x));
      split_store(Bhi, Blo, 4 * cq + 1, rq, make_float4(x[0].y, x[1].y, x[2].y, x[3].y));
      split_store(Bhi, Blo, 4 * cq + 2, rq, make_float4(x[0].z, x[1].z, x[2].z, x[3].z));
      split_store(Bhi, Blo, 4 * cq + 3, rq, make_float4(x[0].w, x[1].w, x[2].w, x[3].w));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t dah = desc_k(smem_u32(Ahi) + 32 * ks), dal = desc_k(smem_u32(Alo) + 32 * ks);
        const uint64_t dbh = desc_k(smem_u32(Bhi) + 32 * ks), dbl = desc_k(smem_u32(Blo) + 32 * ks);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(dal), "l"(dbh), "r"(kIdescKT), "r"(ks > 0 ? 1u : 0u) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(dah), "l"(dbl), "r"(kIdescKT), "r"(1u) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(dah), "l"(dbh), "r"(kIdescKT), "r"(1u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
    if (!done && tid == 0) atomicExch(err, 1);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
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
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] += __uint_as_float(r[k]);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) W[k * 128 + tid] = acc[k];
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

static void run_kt() {
  for (int R : {32, 64, 160, 512, 2048}) {
    const int lda = 256, ldv = 64;
    std::vector<float> A((size_t)R * lda), V((size_t)R * ldv), W(32 * 128);
    srand(3);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 0.01f;
    for (auto& x : V) x = (rand() / (float)RAND_MAX - 0.5f) * 0.01f;
    for (int i = 0; i < 32; ++i) V[(size_t)i * ldv + i] = 1.f;
    float *dA, *dV, *dW; int* derr;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dV, V.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&derr, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, V.data(), V.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(derr, 0, 4); cudaMemset(dW, 0xFF, W.size() * 4);
    const size_t smem = 2 * 16384 + 2 * 4096 + 1024 + 64;
    cudaFuncSetAttribute(probe_kernel_kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel_kt<<<1, 128, smem>>>(dA, lda, dV, ldv, dW, R, derr);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaMemcpy(W.data(), dW, W.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int k = 0; k < 32; ++k)
      for (int c = 0; c < 128; ++c) {
        double s2 = 0;
        for (int i = 0; i < R; ++i) s2 += (double)V[(size_t)i * ldv + k] * A[(size_t)i * lda + c];
        maxerr = fmax(maxerr, fabs(s2 - W[k * 128 + c])); maxref = fmax(maxref, fabs(s2));
      }
    printf("transposed staging, K-major (R = %4d): cuda %s, max err %.3e (max |W| %.3e)\n", R, cudaGetErrorString(e), maxerr, maxref);
    cudaFree(dA); cudaFree(dV); cudaFree(dW); cudaFree(derr);
    if (e != cudaSuccess) break;
  }
}

static void run_control() {
  std::vector<float> A(128 * 32), B(128 * 32), P(128 * 128);
  srand(2);
  for (auto& x : A) x = rand() / (float)RAND_MAX - 0.5f;
  for (auto& x : B) x = rand() / (float)RAND_MAX - 0.5f;
  float *dA, *dB, *dP; int* derr;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dP, P.size() * 4); cudaMalloc(&derr, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(derr, 0, 4); cudaMemset(dP, 0xFF, P.size() * 4);
  const size_t smem = 4 * 16384 + 1024 + 64;
  cudaFuncSetAttribute(control_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  control_kernel<<<1, 128, smem>>>(dA, dB, dP, derr);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      double s = 0;
      for (int k = 0; k < 32; ++k) s += (double)A[i * 32 + k] * B[j * 32 + k];
      maxerr = fmax(maxerr, fabs(s - P[i * 128 + j])); maxref = fmax(maxref, fabs(s));
    }
  printf("control (K-major, M=N=128): cuda %s, max err %.3e (max |P| %.3e), P[0][0..3] = %g %g %g %g\n", cudaGetErrorString(e),
         maxerr, maxref, P[0], P[1], P[2], P[3]);
}

int main() {
  run_control();
  run_kt();
  struct Var { const char* name; uint32_t idesc, lbo, sbo, layout; };
  const uint32_t base = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const Var vars[] = {
      {"MN/MN sw128 lbo4096 sbo1024 (as in the patch)", base | (1u << 15) | (1u << 16), 4096, 1024, 2},
      {"MN/MN sw128 lbo1024 sbo4096 (swapped)", base | (1u << 15) | (1u << 16), 1024, 4096, 2},
      {"K/K idesc on the same tiles", base, 4096, 1024, 2},
      {"A MN, B K", base | (1u << 15), 4096, 1024, 2},
      {"A K, B MN", base | (1u << 16), 4096, 1024, 2},
      {"MN/MN layout 1 (128B_BASE32B)", base | (1u << 15) | (1u << 16), 4096, 1024, 1},
      {"MN/MN layout 0 (no swizzle)", base | (1u << 15) | (1u << 16), 4096, 1024, 0},
  };
  const int R = 32, lda = 128, ldv = 32;
  std::vector<float> A((size_t)R * lda), V((size_t)R * ldv), W(32 * 128);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 0.01f;
  for (auto& x : V) x = (rand() / (float)RAND_MAX - 0.5f) * 0.01f;
  for (int i = 0; i < 32; ++i) V[(size_t)i * ldv + i] = 1.f;
  float *dA, *dV, *dW; int* derr;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dV, V.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&derr, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dV, V.data(), V.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * 16384 + 2 * 4096 + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (const Var& v : vars) {
    cudaMemcpyToSymbol(g_idesc, &v.idesc, 4); cudaMemcpyToSymbol(g_lbo, &v.lbo, 4);
    cudaMemcpyToSymbol(g_sbo, &v.sbo, 4); cudaMemcpyToSymbol(g_layout, &v.layout, 4);
    cudaMemset(derr, 0, 4); cudaMemset(dW, 0xFF, W.size() * 4);
    probe_kernel<<<1, 128, smem>>>(dA, lda, dV, ldv, dW, R, derr);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaMemcpy(W.data(), dW, W.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0, maxgpu = 0;
    for (int k = 0; k < 32; ++k)
      for (int c = 0; c < 128; ++c) {
        double s2 = 0;
        for (int i = 0; i < R; ++i) s2 += (double)V[(size_t)i * ldv + k] * A[(size_t)i * lda + c];
        maxerr = fmax(maxerr, fabs(s2 - W[k * 128 + c])); maxref = fmax(maxref, fabs(s2)); maxgpu = fmax(maxgpu, fabs(W[k * 128 + c]));
      }
    printf("%-48s: cuda %s, max err %.3e, max |ref| %.3e, max |gpu| %.3e\n", v.name, cudaGetErrorString(e), maxerr, maxref, maxgpu);
    if (e != cudaSuccess) break;
  }
  return 0;
}

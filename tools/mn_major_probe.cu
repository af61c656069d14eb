// Probe: tcgen05.mma kind::tf32 with MN-major (transposed) A and B operands in the SWIZZLE_128B canonical
// layout -- D[m][n] = sum_k A[k][m] * B[k][n] with A, B row-major [K][128] in global memory (the natural
// layout of row slabs of a row-major matrix).  Prints the max error against a host reference that
// truncates the operands to TF32.   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_build/mn_probe tools/mn_major_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// MN-major SW128: 32 MN-elements (128 B) contiguous per K row, 8 K rows per 1 KB atom (SBO = 1024 between
// groups of 8 K rows), LBO between 32-element MN blocks
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t ltype = 2) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)ltype << 61);
}
constexpr uint32_t kIdescMN = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

constexpr int K = 32;
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, uint32_t lbo, uint32_t sbo, uint32_t kstep,
                                             uint32_t idesc, int cbstride) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* As = base;           // 4 column blocks x (32 K rows x 128 B)
  unsigned char* Bs = base + 16384;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(base + 32768);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + 32768 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // element (k, c): column block cb = c / 32, 16-byte chunk q = (c % 32) / 4
  if (cbstride == 0) {  // control: K-major staging (row = MN index, 128 B of K), the layout the production kernels use
    for (int idx = tid; idx < 128 * K; idx += 128) {
      const int mn = idx & 127, k = idx >> 7;
      const int off = (mn >> 3) * 1024 + (mn & 7) * 128 + (((k >> 2) ^ (mn & 7)) << 4) + (k & 3) * 4;
      *reinterpret_cast<float*>(As + off) = A[k * 128 + mn];
      *reinterpret_cast<float*>(Bs + off) = B[k * 128 + mn];
    }
  } else if (cbstride < 0) {  // SWIZZLE_128B_BASE32B: 4 K rows x 128 B atoms, 32-byte chunks XORed with (k & 3)
    for (int idx = tid; idx < K * 32; idx += 128) {
      const int k = idx >> 5, c4 = idx & 31, cb = c4 >> 3, q = c4 & 7;  // q: 16-byte chunk of the 128 B row
      const int off = cb * (-cbstride) + (k >> 2) * 512 + (k & 3) * 128 + ((((q >> 1) ^ (k & 3)) << 5) | ((q & 1) << 4));
      *reinterpret_cast<float4*>(As + off) = *reinterpret_cast<const float4*>(A + k * 128 + 4 * c4);
      *reinterpret_cast<float4*>(Bs + off) = *reinterpret_cast<const float4*>(B + k * 128 + 4 * c4);
    }
  } else
  for (int idx = tid; idx < K * 32; idx += 128) {
    const int k = idx >> 5, c4 = idx & 31, cb = c4 >> 3, q = c4 & 7;
    const int off = cb * cbstride + (k >> 3) * (cbstride == 4096 ? 1024 : 4096) + (k & 7) * 128 + ((q ^ (k & 7)) << 4);
    *reinterpret_cast<float4*>(As + off) = *reinterpret_cast<const float4*>(A + k * 128 + 4 * c4);
    *reinterpret_cast<float4*>(Bs + off) = *reinterpret_cast<const float4*>(B + k * 128 + 4 * c4);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t lt = cbstride < 0 ? 1u : 2u;
      const uint64_t da = desc_mn(smem_u32(As) + ks * kstep, lbo, sbo, lt), db = desc_mn(smem_u32(Bs) + ks * kstep, lbo, sbo, lt);
      const uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                   "l"(da), "l"(db), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
  }
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
  if (!done) __trap();
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(32 * warp) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[tid * 128 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }


int main() {
  std::vector<float> A(K * 128), B(K * 128), D(128 * 128);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  struct Var { const char* name; uint32_t lbo, sbo, kstep; int cbstride; };
  // layout a: column block cb at cb * 4096, K rows contiguous inside it (8-row groups 1 KB apart)
  // layout b: 8-row K group g at g * 4096, column blocks 1 KB apart inside it
  const Var vars[] = {{"control K-major", 16, 1024, 32, 0}, {"base32b lbo4096 sbo512 k1024", 4096, 512, 1024, -4096},
                      {"base32b lbo512 sbo4096 k1024", 512, 4096, 1024, -4096}, {"a lbo4096 sbo1024 k1024", 4096, 1024, 1024, 4096}, {"a lbo1024 sbo4096 k1024", 1024, 4096, 1024, 4096},
                      {"b lbo1024 sbo4096 k4096", 1024, 4096, 4096, 1024}, {"b lbo4096 sbo1024 k4096", 4096, 1024, 4096, 1024}};
  for (int t = 0; t < 4; t += 3) {
    srand(1);
    for (int k = 0; k < K; ++k)
      for (int c = 0; c < 128; ++c) {
        float a, b;
        if (t == 0) { a = (k == 0); b = (k == 0) ? c : 0; }             // D[m][n] = n
        else if (t == 1) { a = (k == 0) ? c : 0; b = (k == 0); }        // D[m][n] = m
        else if (t == 2) { a = 1; b = (k == 13) ? 1 : 0; }              // D = 1
        else { a = (float)rand() / RAND_MAX - 0.5f; b = (float)rand() / RAND_MAX - 0.5f; }
        A[k * 128 + c] = a; B[k * 128 + c] = b;
      }
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    for (const Var& v : vars) {
      cudaMemset(dD, 0, D.size() * 4);
      probe<<<1, 128, 40000>>>(dA, dB, dD, v.lbo, v.sbo, v.kstep, v.cbstride == 0 ? (kIdescMN & ~((1u << 15) | (1u << 16))) : kIdescMN, v.cbstride);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { printf("MN_PROBE launch config error: %s\n", cudaGetErrorString(e)); return 1; }
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("MN_PROBE launch error: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
          double s = 0;
          for (int k = 0; k < K; ++k) s += (double)trunc_tf32(A[k * 128 + m]) * trunc_tf32(B[k * 128 + n]);
          maxerr = fmax(maxerr, fabs(s - D[m * 128 + n]));
          maxref = fmax(maxref, fabs(s));
        }
      printf("test %d var [%s]: maxerr %.3e maxref %.3e %s | D[0][0..5] %g %g %g %g %g %g D[0][32..33] %g %g D[1][0] %g D[5][0] %g D[33][0] %g D[64][7] %g\n",
             t, v.name, maxerr, maxref, maxerr < 1e-5 * maxref + 1e-6 ? "OK" : "MISMATCH", D[0], D[1], D[2], D[3], D[4], D[5], D[32],
             D[33], D[128], D[5 * 128], D[33 * 128], D[64 * 128 + 7]);
    }
  }
  return 0;
}

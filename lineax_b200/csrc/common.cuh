// Shared device helpers for the lineax_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <limits>

#include "../../include/lineax_b200.h"

namespace lxb {

constexpr int kNumSMs = 148;  // B200
constexpr unsigned kFull = 0xffffffffu;

extern std::atomic<int64_t> g_launch_count;
inline void count_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }

#define LXB_CUDA_CHECK_LAUNCH()               \
  do {                                        \
    ::lxb::count_launch();                    \
    cudaError_t e__ = cudaPeekAtLastError();  \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

#define LXB_CUDA_TRY(expr)                    \
  do {                                        \
    cudaError_t e__ = (expr);                 \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

template <typename T>
struct Num;
template <>
struct Num<float> {
  __host__ __device__ static float eps() { return 1.1920928955078125e-07f; }
  __host__ __device__ static float inf() { return __builtin_huge_valf(); }
  __host__ __device__ static float nan() { return __builtin_nanf(""); }
  __host__ __device__ static float max() { return 3.40282346638528859812e+38f; }
};
template <>
struct Num<double> {
  __host__ __device__ static double eps() { return 2.220446049250313e-16; }
  __host__ __device__ static double inf() { return __builtin_huge_val(); }
  __host__ __device__ static double nan() { return __builtin_nan(""); }
  __host__ __device__ static double max() { return 1.79769313486231570815e+308; }
};

__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float abs_(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_(double a) { return fabs(a); }
__device__ __forceinline__ float sqrt_(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ bool finite_(float a) { return isfinite(a); }
__device__ __forceinline__ bool finite_(double a) { return isfinite(a); }

// ---- warp reductions ------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// NaN-propagating max of non-negative values (jnp.max semantics, lineax/_norm.py:130).
// Values are |x| >= 0 or NaN; comparing the raw bit patterns as unsigned integers
// orders all non-negative floats and puts every NaN above +inf, so a NaN wins.
__device__ __forceinline__ float warp_absmax(float v) {
  unsigned u = __float_as_uint(v) & 0x7fffffffu;
  u = __reduce_max_sync(kFull, u);
  return __uint_as_float(u);
}
__device__ __forceinline__ double warp_absmax(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(kFull, u, o);
    u = w > u ? w : u;
  }
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ float absmax2(float a, float b) {
  unsigned x = __float_as_uint(a) & 0x7fffffffu, y = __float_as_uint(b) & 0x7fffffffu;
  return __uint_as_float(x > y ? x : y);
}
__device__ __forceinline__ double absmax2(double a, double b) {
  unsigned long long x = (unsigned long long)__double_as_longlong(a) & 0x7fffffffffffffffull;
  unsigned long long y = (unsigned long long)__double_as_longlong(b) & 0x7fffffffffffffffull;
  return __longlong_as_double((long long)(x > y ? x : y));
}

// ---- block reductions (all threads get the result) -------------------------
// `scratch` must hold >= 32 * K elements of T. Deterministic order.
template <typename T, int K>
__device__ __forceinline__ void block_sum(T (&v)[K], T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // protect scratch reuse
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < K; ++i) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < K; ++i) {
    T t = lane < nw ? scratch[i * 32 + lane] : T(0);
    v[i] = warp_sum(t);
  }
}

template <typename T, int K>
__device__ __forceinline__ void block_absmax(T (&v)[K], T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) v[i] = warp_absmax(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < K; ++i) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < K; ++i) {
    T t = lane < nw ? scratch[i * 32 + lane] : T(0);
    v[i] = warp_absmax(t);
  }
}

// ---- cp.async (LDGSTS) helpers ------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// streaming 128-bit global load that does not pollute L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace lxb

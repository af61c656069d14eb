// CG with the operator resident ON CHIP for the whole solve: BASELINE configs[2]
// (4096 x 256x256 SPD fp32).  A 256x256 fp32 matrix is 256 KB > 227 KB of shared memory, so one
// CTA (512 threads) keeps rows 0..127 in REGISTERS (64 per thread) and rows 128..255 in shared
// memory (128 KB): A is read from HBM exactly once per SOLVE instead of once per iteration
// (lineax/_solver/cg.py:114-227 reads it 1 + k + k/10 times).  Same arithmetic statements as
// cg.cu; only the matvec's summation order differs.
#include "krylov_cta.cuh"
#include "cg_resident.cuh"

namespace lxb {

constexpr int kResN = 256;
constexpr int kResThreads = 512;
constexpr int kResWarps = kResThreads / 32;   // 16
constexpr int kRowsPerWarp = 8;               // per half (register half / shared half)

// Reduce 16 per-lane partial row sums across the warp: on exit lane l holds the full sum of
// row (l >> 1) (both lanes of a pair hold the same value).  15 + 1 shuffles instead of 80.
__device__ __forceinline__ float warp_transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(kFull, send, bit);
    }
  }
  return v[0] + __shfl_xor_sync(kFull, v[0], 1);
}

__global__ void __launch_bounds__(kResThreads, 1) cg_resident_kernel(KrylovParams<float> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int n = kResN;
  float* sAh = reinterpret_cast<float*>(smem_raw);  // rows 128..255, 128 x 256
  float* sb = sAh + 128 * n;
  float* sy = sb + n;
  float* sp = sy + n;
  float* sq = sp + n;
  float* wpart = sq + n;  // 64 per-warp partials
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool has_scale = !(p.rtol == 0.f && p.atol == 0.f);
  const float sign = (p.flags & LXB_NSD) ? -1.f : 1.f;
  const float rcond = 2.f * Num<float>::eps() * float(n);
  float4 areg[kRowsPerWarp][2];  // rows 8*warp .. 8*warp+7, column chunks lane and lane+32

  // q = sign * A x for a shared-memory vector x, written to sq; each warp also leaves its partial
  // of <q, x> in wpart[warp] when WITH_DOT.  Ends with one __syncthreads().
  auto matvec = [&](const float* x, bool with_dot) {
    const float4 x0 = reinterpret_cast<const float4*>(x)[lane];
    const float4 x1 = reinterpret_cast<const float4*>(x)[lane + 32];
    float acc[16];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const float4 a0 = areg[r][0], a1 = areg[r][1];
      float s = a0.x * x0.x;
      s = fmaf(a0.y, x0.y, s); s = fmaf(a0.z, x0.z, s); s = fmaf(a0.w, x0.w, s);
      s = fmaf(a1.x, x1.x, s); s = fmaf(a1.y, x1.y, s); s = fmaf(a1.z, x1.z, s); s = fmaf(a1.w, x1.w, s);
      acc[r] = s;
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const float4* row = reinterpret_cast<const float4*>(sAh + (size_t)(warp * kRowsPerWarp + r) * n);
      const float4 a0 = row[lane], a1 = row[lane + 32];
      float s = a0.x * x0.x;
      s = fmaf(a0.y, x0.y, s); s = fmaf(a0.z, x0.z, s); s = fmaf(a0.w, x0.w, s);
      s = fmaf(a1.x, x1.x, s); s = fmaf(a1.y, x1.y, s); s = fmaf(a1.z, x1.z, s); s = fmaf(a1.w, x1.w, s);
      acc[8 + r] = s;
    }
    const float tot = sign * warp_transpose_reduce16(acc, lane);
    const int rr = lane >> 1;  // 0..15: 0..7 register rows, 8..15 shared rows
    const int row = rr < 8 ? warp * kRowsPerWarp + rr : 128 + warp * kRowsPerWarp + (rr - 8);
    float contrib = 0.f;
    if ((lane & 1) == 0) {
      sq[row] = tot;
      contrib = tot * x[row];
    }
    if (with_dot) {
      contrib = warp_sum(contrib);
      if (lane == 0) wpart[warp] = contrib;
    }
    __syncthreads();
  };
  // fixed-order sum of the 16 per-warp partials (same bits in every thread)
  auto sum_wpart = [&]() { return warp_sum(lane < kResWarps ? wpart[lane] : 0.f); };

  for (int64_t sys = blockIdx.x; sys < p.batch; sys += gridDim.x) {
    const float* Ag = p.A + sys * p.sA;
    // stage: upper half of the rows straight into registers, lower half into shared memory
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const float4* row = reinterpret_cast<const float4*>(Ag + (size_t)(warp * kRowsPerWarp + r) * n);
      areg[r][0] = ldg_stream(row + lane);
      areg[r][1] = ldg_stream(row + lane + 32);
    }
    {
      const float* src = Ag + (size_t)128 * n;
      for (int c = tid; c < 128 * n / 4; c += kResThreads) cp_async16(sAh + (size_t)c * 4, src + (size_t)c * 4);
      cp_async_commit();
    }
    // every thread tid < n owns element tid of b, y, r, p, diff in registers; p (and y when a
    // true-residual step needs it) are mirrored in shared memory for the matvec
    const bool own = tid < n;
    float bi = 0.f, yi = 0.f, ri = 0.f, pi = 0.f, di = Num<float>::inf();
    if (own) {
      bi = p.b[sys * p.sb + tid];
      yi = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + tid] : 0.f;
      sy[tid] = yi;
    }
    cp_async_wait<0>();
    __syncthreads();
    matvec(sy, false);  // r0 = b - A y0 (always evaluated, cg.py:128)
    float gamma, norm1 = 0.f, norm2 = 0.f;
    // reduction of (sum r_i p_i | r_i r_i, max |r/b_scale|, max |diff/y_scale|) over the 8 owner warps
    auto reduce3 = [&](float sval) {
      const float bs = p.atol + p.rtol * fabsf(bi);
      const float ys = p.atol + p.rtol * fabsf(yi);
      float m1 = own ? ri / bs : 0.f, m2 = own ? di / ys : 0.f;
      sval = warp_sum(own ? sval : 0.f);
      m1 = warp_absmax(m1);
      m2 = warp_absmax(m2);
      if (lane == 0) {
        wpart[16 + warp] = sval;
        wpart[32 + warp] = m1;
        wpart[48 + warp] = m2;
      }
      __syncthreads();
      // every warp folds the 16 per-warp partials with shuffles (same order => same bits everywhere)
      const bool in = lane < kResWarps;
      norm1 = warp_absmax(in ? wpart[32 + lane] : 0.f);
      norm2 = warp_absmax(in ? wpart[48 + lane] : 0.f);
      return warp_sum(in ? wpart[16 + lane] : 0.f);
    };
    if (own) {
      ri = bi - sq[tid];
      pi = ri;
      sp[tid] = pi;
    }
    gamma = reduce3(pi * ri);  // barrier inside: sp visible
    int64_t step = 0;
    while (true) {
      // cond_fun, cg.py:162-167 (norms from the previous body / initial state, diff = inf at first)
      if (!(gamma > 0.f)) break;
      if (!(step < p.max_steps)) break;
      if (has_scale && !((norm1 > 1.f) || (norm2 > 1.f))) break;
      matvec(sp, true);
      const float ip = sum_wpart();
      float alpha = gamma / ip;
      if (!(fabsf(ip) > 100.f * rcond * fabsf(gamma))) alpha = Num<float>::nan();
      step += 1;
      const bool stable = p.stabilise_every == 1 ||
                          (p.stabilise_every > 1 && (step % p.stabilise_every) == 0);
      if (own) {
        di = alpha * pi;
        yi = yi + di;
        if (!stable) ri = ri - alpha * sq[tid];
      }
      if (stable) {  // cg.py:187-200: r = b - A y
        if (own) sy[tid] = yi;
        __syncthreads();
        matvec(sy, false);
        if (own) ri = bi - sq[tid];
      }
      const float gn = reduce3(ri * ri);
      const float beta = gn / gamma;
      gamma = gn;
      if (own) {
        pi = ri + beta * pi;
        sp[tid] = pi;
      }
      __syncthreads();
    }
    if (own) p.x[sys * n + tid] = (p.flags & LXB_NSD) ? -yi : yi;
    if (tid == 0) {
      p.result[sys] = krylov_final_result(step, p.max_steps, p.flags, has_scale);
      p.num_steps[sys] = (int32_t)step;
    }
    __syncthreads();
  }
}

bool cg_resident_applicable(const KrylovParams<float>& p) {
  return p.n == kResN && p.M == nullptr && aligned16(p.A) && (p.sA % 4 == 0) && p.batch >= kNumSMs / 2;
}

int cg_resident_launch(KrylovParams<float> p, cudaStream_t st) {
  const size_t smem = (size_t)(128 * kResN + 5 * kResN + 64) * sizeof(float);
  auto kern = cg_resident_kernel;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = p.batch < kNumSMs ? p.batch : kNumSMs;
  kern<<<(unsigned)blocks, kResThreads, smem, st>>>(p);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

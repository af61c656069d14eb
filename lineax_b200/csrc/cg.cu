// Preconditioned conjugate gradients, lineax/_solver/cg.py:114-227, as ONE persistent kernel:
// matvec, dots, axpys, the max-norm convergence test, the rcond guard and the
// stabilise_every true-residual recompute all run inside the kernel, each system exits
// as soon as its own test passes (observably equal to JAX's masked lock-step loop,
// SURVEY.md App. B-3).
#include <type_traits>

#include "cg_resident.cuh"
#include "krylov_cta.cuh"
#include "krylov_grid_api.cuh"

namespace lxb {

template <typename T>
__global__ void __launch_bounds__(kKrylovThreads) cg_cta_kernel(KrylovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int npad = (n + 3) & ~3;  // keep every vector 16-byte aligned
  T* sb = reinterpret_cast<T*>(smem_raw);
  T* sy = sb + npad;
  T* sr = sy + npad;
  T* sp = sr + npad;
  T* sq = sp + npad;   // A p, then z = M r
  T* sd = sq + npad;   // diff
  T* red = sd + npad;  // 96 elements of reduction scratch
  T* sA = red + 96;    // staged matrix (a_smem only)
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const T sign = (p.flags & LXB_NSD) ? T(-1) : T(1);  // cg.py:100-101: operator = -operator
  const T rcond = T(2) * Num<T>::eps() * T(n);        // cg.py:131, _misc.py:30-38

  for (int64_t sys = blockIdx.x; sys < p.batch; sys += gridDim.x) {
    const T* Ag = p.A + sys * p.sA;
    const T* A = Ag;
    if (p.a_smem) {
      cta_stage_matrix<T>(Ag, sA, (size_t)n * n);
      A = sA;
    }
    const T* Mg = p.M ? p.M + sys * p.sM : nullptr;
    for (int i = tid; i < n; i += nt) {
      sb[i] = p.b[sys * p.sb + i];
      sy[i] = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + i] : T(0);
    }
    __syncthreads();
    // r0 = b - A y0 (cg.py:128: always evaluated, so non-finite A poisons r0 even for y0 = 0)
    cta_matvec<T>(A, n, n, n, sy, sq, sign);
    for (int i = tid; i < n; i += nt) sr[i] = sb[i] - sq[i];
    __syncthreads();
    if (Mg) {
      cta_matvec<T>(Mg, n, n, n, sr, sp, T(1));
    } else {
      for (int i = tid; i < n; i += nt) sp[i] = sr[i];
      __syncthreads();
    }
    T g[1] = {T(0)};
    for (int i = tid; i < n; i += nt) g[0] = fma_(sp[i], sr[i], g[0]);
    block_sum<T, 1>(g, red);
    T gamma = g[0];
    int64_t step = 0;
    bool diff_inf = true;

    while (true) {
      // cond_fun, cg.py:162-167
      if (!(gamma > T(0))) break;
      if (!(step < p.max_steps)) break;
      if (!cta_not_converged<T>(sr, sd, sy, sb, n, p.rtol, p.atol, has_scale, diff_inf, red + 32))
        break;
      // body_fun, cg.py:169-207
      cta_matvec<T>(A, n, n, n, sp, sq, sign);
      T ip[1] = {T(0)};
      for (int i = tid; i < n; i += nt) ip[0] = fma_(sq[i], sp[i], ip[0]);
      block_sum<T, 1>(ip, red);
      T alpha = gamma / ip[0];
      if (!(abs_(ip[0]) > T(100) * rcond * abs_(gamma))) alpha = Num<T>::nan();  // cg.py:174-178
      step += 1;
      const bool stable = p.stabilise_every == 1 ||
                          (p.stabilise_every > 1 && (step % p.stabilise_every) == 0);
      for (int i = tid; i < n; i += nt) {
        const T d = alpha * sp[i];
        sd[i] = d;
        sy[i] = sy[i] + d;
        if (!stable) sr[i] = sr[i] - alpha * sq[i];
      }
      diff_inf = false;
      __syncthreads();
      if (stable) {  // cg.py:187-200: r = b - A y
        cta_matvec<T>(A, n, n, n, sy, sq, sign);
        for (int i = tid; i < n; i += nt) sr[i] = sb[i] - sq[i];
        __syncthreads();
      }
      const T* z = sr;
      if (Mg) {
        cta_matvec<T>(Mg, n, n, n, sr, sq, T(1));
        z = sq;
      }
      T gn[1] = {T(0)};
      for (int i = tid; i < n; i += nt) gn[0] = fma_(z[i], sr[i], gn[0]);
      block_sum<T, 1>(gn, red);
      const T beta = gn[0] / gamma;
      gamma = gn[0];
      for (int i = tid; i < n; i += nt) sp[i] = z[i] + beta * sp[i];
      __syncthreads();
    }
    for (int i = tid; i < n; i += nt) p.x[sys * n + i] = (p.flags & LXB_NSD) ? -sy[i] : sy[i];
    if (tid == 0) {
      p.result[sys] = krylov_final_result(step, p.max_steps, p.flags, has_scale);
      p.num_steps[sys] = (int32_t)step;
    }
    __syncthreads();
  }
}

constexpr size_t kMaxSmemK = 227 * 1024;

template <typename T>
int cg_dispatch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (p.batch < 0 || p.n < 0 || !p.A || !p.b || !p.x || !p.result || !p.num_steps) return LXB_E_BADARG;
  if (p.batch == 0) return 0;
  const size_t npad = ((size_t)p.n + 3) & ~(size_t)3;
  const size_t vec_bytes = (6 * npad + 96) * sizeof(T);
  if (use_grid_tier(p.batch, p.n, p.n) || vec_bytes > kMaxSmemK) return cg_grid_launch<T>(p, ws, ws_bytes, st);
  if constexpr (std::is_same<T, float>::value) {
    // 256x256 fp32: operator resident in registers + shared memory for the whole solve
    if (cg_resident_applicable(p) && !getenv("LXB_NO_RESIDENT")) return cg_resident_launch(p, st);
  }
  const size_t mat_bytes = (size_t)p.n * p.n * sizeof(T);
  p.a_smem = (vec_bytes + mat_bytes <= kMaxSmemK) && p.n > 0;
  const size_t smem = vec_bytes + (p.a_smem ? mat_bytes : 0);
  auto kern = cg_cta_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kKrylovThreads, smem));
  if (occ < 1) occ = 1;
  int64_t cap = (int64_t)kNumSMs * occ;
  if (!p.a_smem) cap = l2_resident_cap(cap, mat_bytes);
  const int64_t blocks = p.batch < cap ? p.batch : cap;
  kern<<<(unsigned)blocks, kKrylovThreads, smem, st>>>(p);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_CG(sfx, T)                                                                         \
  extern "C" int lxb_cg_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b,          \
                              const T* Minv, int64_t stride_M, T* x, int32_t* result,              \
                              int32_t* num_steps, int64_t batch, int32_t n, T rtol, T atol,        \
                              int32_t max_steps, int32_t stabilise_every, int32_t flags,           \
                              void* workspace, size_t workspace_bytes, lxb_stream_t stream) {      \
    lxb::KrylovParams<T> p{};                                                                      \
    p.A = A; p.sA = stride_A; p.b = b; p.sb = stride_b; p.M = Minv; p.sM = stride_M; p.x = x;      \
    p.result = result; p.num_steps = num_steps; p.batch = batch; p.m = n; p.n = n;                 \
    p.rtol = rtol; p.atol = atol; p.max_steps = max_steps; p.stabilise_every = stabilise_every;    \
    p.flags = flags;                                                                               \
    return lxb::cg_dispatch<T>(p, workspace, workspace_bytes, (cudaStream_t)stream);                                           \
  }                                                                                                \
  extern "C" size_t lxb_cg_workspace_##sfx(int64_t batch, int32_t n) {                             \
    const bool big = (6 * (((size_t)n + 3) & ~(size_t)3) + 96) * sizeof(T) > lxb::kMaxSmemK;       \
    return (lxb::use_grid_tier(batch, n, n) || big) ? lxb::cg_grid_ws_bytes<T>(n) : 0;             \
  }
LXB_DEF_CG(f32, float)
LXB_DEF_CG(f64, double)

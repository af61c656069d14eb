// Right-preconditioned BiCGStab, lineax/_solver/bicgstab.py:78-205, as one persistent kernel
// (one CTA per system; two matvecs, five dots, the axpys and the breakdown / convergence
// tests per iteration, no host round trips).
#include "krylov_cta.cuh"
#include "krylov_grid_api.cuh"

namespace lxb {

template <typename T>
__device__ __forceinline__ bool bicg_breakdown(T omega, T alpha, T rho, bool x64) {
  // bicgstab.py:107-113: `== 0` under jax_enable_x64, otherwise the SIGNED `< 1e-16` test
  if (x64) return omega == T(0) || alpha == T(0) || rho == T(0);
  const T t = T(1e-16);
  return omega < t || alpha < t || rho < t;
}

template <typename T>
__global__ void __launch_bounds__(kKrylovThreads) bicgstab_cta_kernel(KrylovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int npad = (n + 3) & ~3;
  T* sb = reinterpret_cast<T*>(smem_raw);
  T* sy = sb + npad;
  T* sr0 = sy + npad;
  T* sr = sr0 + npad;
  T* sp = sr + npad;
  T* sv = sp + npad;
  T* ss = sv + npad;
  T* st = ss + npad;
  T* sx = st + npad;  // x_hat = M p
  T* sz = sx + npad;  // z = M s
  T* sd = sz + npad;  // diff
  T* red = sd + npad;
  T* sA = red + 96;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const bool x64 = (p.flags & LXB_X64_BREAKDOWN) != 0;

  for (int64_t sys = blockIdx.x; sys < p.batch; sys += gridDim.x) {
    const T* A = p.A + sys * p.sA;
    if (p.a_smem) {
      cta_stage_matrix<T>(A, sA, (size_t)n * n);
      A = sA;
    }
    const T* Mg = p.M ? p.M + sys * p.sM : nullptr;
    for (int i = tid; i < n; i += nt) {
      sb[i] = p.b[sys * p.sb + i];
      sy[i] = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + i] : T(0);
      sp[i] = T(0);
      sv[i] = T(0);
    }
    __syncthreads();
    cta_matvec<T>(A, n, n, n, sy, st, T(1));
    for (int i = tid; i < n; i += nt) {
      const T r = sb[i] - st[i];
      sr0[i] = r;
      sr[i] = r;
    }
    __syncthreads();
    T alpha = T(1), omega = T(1), rho = T(1);
    int64_t step = 0;
    bool diff_inf = true;
    while (true) {
      if (bicg_breakdown(omega, alpha, rho, x64)) break;
      if (!cta_not_converged<T>(sr, sd, sy, sb, n, p.rtol, p.atol, has_scale, diff_inf, red + 32)) break;
      if (!(step < p.max_steps)) break;
      T d1[1] = {T(0)};
      for (int i = tid; i < n; i += nt) d1[0] = fma_(sr0[i], sr[i], d1[0]);
      block_sum<T, 1>(d1, red);
      const T rho_new = d1[0];
      const T beta = (rho_new / rho) * (alpha / omega);
      for (int i = tid; i < n; i += nt) sp[i] = sr[i] + beta * (sp[i] - omega * sv[i]);
      __syncthreads();
      const T* xh = sp;
      if (Mg) {
        cta_matvec<T>(Mg, n, n, n, sp, sx, T(1));
        xh = sx;
      }
      cta_matvec<T>(A, n, n, n, xh, sv, T(1));
      T d2[1] = {T(0)};
      for (int i = tid; i < n; i += nt) d2[0] = fma_(sr0[i], sv[i], d2[0]);
      block_sum<T, 1>(d2, red);
      alpha = rho_new / d2[0];
      for (int i = tid; i < n; i += nt) ss[i] = sr[i] - alpha * sv[i];
      __syncthreads();
      const T* z = ss;
      if (Mg) {
        cta_matvec<T>(Mg, n, n, n, ss, sz, T(1));
        z = sz;
      }
      cta_matvec<T>(A, n, n, n, z, st, T(1));
      T d3[2] = {T(0), T(0)};
      for (int i = tid; i < n; i += nt) {
        d3[0] = fma_(ss[i], st[i], d3[0]);
        d3[1] = fma_(st[i], st[i], d3[1]);
      }
      block_sum<T, 2>(d3, red);
      omega = d3[0] / d3[1];
      for (int i = tid; i < n; i += nt) {
        const T d = alpha * xh[i] + omega * z[i];
        sd[i] = d;
        sy[i] = sy[i] + d;
        sr[i] = ss[i] - omega * st[i];
      }
      diff_inf = false;
      rho = rho_new;
      step += 1;
      __syncthreads();
    }
    int result = krylov_final_result(step, p.max_steps, p.flags, has_scale);
    // bicgstab.py:199-202: breakdown only matters if we did not converge
    const bool nc = cta_not_converged<T>(sr, sd, sy, sb, n, p.rtol, p.atol, has_scale, diff_inf, red + 32);
    if (bicg_breakdown(omega, alpha, rho, x64) && nc) result = LXB_BREAKDOWN;
    for (int i = tid; i < n; i += nt) p.x[sys * n + i] = sy[i];
    if (tid == 0) {
      p.result[sys] = result;
      p.num_steps[sys] = (int32_t)step;
    }
    __syncthreads();
  }
}

template <typename T>
int bicgstab_dispatch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (p.batch < 0 || p.n < 0 || !p.A || !p.b || !p.x || !p.result || !p.num_steps) return LXB_E_BADARG;
  if (p.batch == 0) return 0;
  const size_t kMax = 227 * 1024;
  const size_t npad = ((size_t)p.n + 3) & ~(size_t)3;
  const size_t vec_bytes = (11 * npad + 96) * sizeof(T);
  if (use_grid_tier(p.batch, p.n, p.n) || vec_bytes > kMax) return bicgstab_grid_launch<T>(p, ws, ws_bytes, st);
  const size_t mat_bytes = (size_t)p.n * p.n * sizeof(T);
  p.a_smem = (vec_bytes + mat_bytes <= kMax) && p.n > 0;
  const size_t smem = vec_bytes + (p.a_smem ? mat_bytes : 0);
  auto kern = bicgstab_cta_kernel<T>;
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

#define LXB_DEF_BICGSTAB(sfx, T)                                                                   \
  extern "C" int lxb_bicgstab_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b,    \
                                    const T* Minv, int64_t stride_M, T* x, int32_t* result,        \
                                    int32_t* num_steps, int64_t batch, int32_t n, T rtol, T atol,  \
                                    int32_t max_steps, int32_t flags, void* workspace,             \
                                    size_t workspace_bytes, lxb_stream_t stream) {                 \
    lxb::KrylovParams<T> p{};                                                                      \
    p.A = A; p.sA = stride_A; p.b = b; p.sb = stride_b; p.M = Minv; p.sM = stride_M; p.x = x;      \
    p.result = result; p.num_steps = num_steps; p.batch = batch; p.m = n; p.n = n;                 \
    p.rtol = rtol; p.atol = atol; p.max_steps = max_steps; p.flags = flags;                        \
    return lxb::bicgstab_dispatch<T>(p, workspace, workspace_bytes, (cudaStream_t)stream);                                     \
  }                                                                                                \
  extern "C" size_t lxb_bicgstab_workspace_##sfx(int64_t batch, int32_t n) {                       \
    const bool big = (11 * (((size_t)n + 3) & ~(size_t)3) + 96) * sizeof(T) > 227 * 1024;          \
    return (lxb::use_grid_tier(batch, n, n) || big) ? lxb::bicgstab_grid_ws_bytes<T>(n) : 0;       \
  }
LXB_DEF_BICGSTAB(f32, float)
LXB_DEF_BICGSTAB(f64, double)

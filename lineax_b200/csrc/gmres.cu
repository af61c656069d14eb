// Restarted GMRES, lineax/_solver/gmres.py:106-413, as one persistent kernel per batch
// (one CTA per system).  Reproduced quirks: `num_steps` counts RESTARTS,
// the first pass is a dummy (gmres.py:195,313-316), all `restart` Arnoldi steps always run
// unless Arnoldi itself breaks down, breakdown is deferred by one outer step, single-pass
// classical Gram-Schmidt against ALL restart+1 basis columns, stagnation counter on the
// max-norm of the preconditioned residual, Hessenberg least squares by Householder QR
// (the reference calls its own QR solver, gmres.py:301-303).
// The Krylov basis lives in shared memory when it fits, otherwise in the caller's workspace
// (L2-resident for the sizes of interest).
#include "krylov_cta.cuh"
#include "krylov_grid_api.cuh"

namespace lxb {

// two_norm with lineax's size-1 shortcut (_norm.py:74-80). All threads get the value.
template <typename T>
__device__ __forceinline__ T cta_two_norm(const T* x, int n, T* red) {
  if (n == 1) return abs_(x[0]);
  T s[1] = {T(0)};
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[0] = fma_(x[i], x[i], s[0]);
  block_sum<T, 1>(s, red);
  return sqrt_(s[0]);
}

// Least squares  min || rhs - Mt z ||  for the (R+1) x R matrix Mt = coeff^T by Householder QR
// (geqr2 + ormqr + trtrs ordering).  `Q` holds Mt row-major with leading dimension R, is
// destroyed; rhs (R+1) is destroyed; z (R) receives the solution.  Block-cooperative.
template <typename T>
__device__ void cta_hessenberg_lstsq(T* Q, T* rhs, T* z, int R, T* sc) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int rows = R + 1;
  for (int j = 0; j < R; ++j) {
    if (tid == 0) {
      // larfg on column j, rows j..rows-1
      const T alpha = Q[j * R + j];
      T ssq = T(0);
      for (int i = j + 1; i < rows; ++i) ssq = fma_(Q[i * R + j], Q[i * R + j], ssq);
      T tau = T(0), beta = alpha;
      if (ssq != T(0)) {
        const T nrm = sqrt_(alpha * alpha + ssq);
        beta = alpha >= T(0) ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        const T scal = T(1) / (alpha - beta);
        for (int i = j + 1; i < rows; ++i) Q[i * R + j] *= scal;
      }
      Q[j * R + j] = beta;
      sc[0] = tau;
    }
    __syncthreads();
    const T tau = sc[0];
    // apply H = I - tau v v^T (v = [1; Q[j+1:, j]]) to the trailing columns and to rhs
    for (int c = j + 1 + tid; c <= R; c += nt) {
      T* col = c < R ? Q + c : rhs;
      const int ld = c < R ? R : 1;
      T dot = col[j * ld];
      for (int i = j + 1; i < rows; ++i) dot = fma_(Q[i * R + j], col[i * ld], dot);
      const T f = tau * dot;
      col[j * ld] -= f;
      for (int i = j + 1; i < rows; ++i) col[i * ld] = fma_(-f, Q[i * R + j], col[i * ld]);
    }
    __syncthreads();
  }
  // R z = (Q^T rhs)[:R]
  if (tid == 0) {
    for (int k = R - 1; k >= 0; --k) {
      T s = rhs[k];
      for (int c = k + 1; c < R; ++c) s = fma_(-Q[k * R + c], z[c], s);
      z[k] = s / Q[k * R + k];
    }
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kKrylovThreads) gmres_cta_kernel(KrylovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, R = p.restart;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int npad = (n + 3) & ~3;
  T* sb = reinterpret_cast<T*>(smem_raw);
  T* sy = sb + npad;
  T* sr = sy + npad;
  T* sw = sr + npad;
  T* sd = sw + npad;                 // diff
  T* st = sd + npad;                 // scratch vector (A v before the preconditioner)
  T* red = st + npad;                // 96
  T* proj = red + 96;                // R + 1
  T* zv = proj + (R + 1);            // R
  T* rhs = zv + R;                   // R + 1
  T* coeff = rhs + (R + 1);          // R x (R + 1), row k = column k of the Hessenberg
  T* Qm = coeff + (size_t)R * (R + 1);  // (R + 1) x R work copy for the QR
  T* sc = Qm + (size_t)R * (R + 1);  // 4
  T* tail = sc + 4;
  tail = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(tail) + 15) & ~(uintptr_t)15);
  T* Vs = tail;                      // basis (R + 1) x npad when p.restart basis fits (flag in ws_stride<0)
  const bool basis_smem = p.ws == nullptr;
  T* sA = basis_smem ? Vs + (size_t)(R + 1) * npad : Vs;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const T eps = Num<T>::eps();

  for (int64_t sys = blockIdx.x; sys < p.batch; sys += gridDim.x) {
    const T* A = p.A + sys * p.sA;
    if (p.a_smem) {
      cta_stage_matrix<T>(A, sA, (size_t)n * n);
      A = sA;
    }
    const T* Mg = p.M ? p.M + sys * p.sM : nullptr;
    T* V = basis_smem ? Vs : p.ws + (size_t)blockIdx.x * p.ws_stride;
    for (int i = tid; i < n; i += nt) {
      sb[i] = p.b[sys * p.sb + i];
      sy[i] = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + i] : T(0);
      sr[i] = T(0);  // dummy residual, gmres.py:195
    }
    __syncthreads();
    bool breakdown = false, deferred = false, diff_inf = true;
    T r_min = Num<T>::inf();
    int64_t step = 0;
    int stag = 0;
    while (true) {
      // cond_fun, gmres.py:143-158
      bool go = !deferred && stag < p.stagnation_iters;
      if (go) go = cta_not_converged<T>(sr, sd, sy, sb, n, p.rtol, p.atol, has_scale, diff_inf, red + 32);
      go = (go && step < p.max_steps) || step == 0;
      if (!go) break;
      bool bd_new = false;
      if (step > 0) {
        // ---- main_gmres, gmres.py:254-310
        const T beta0 = cta_two_norm<T>(sr, n, red);
        const bool init_bd = beta0 < eps;
        const T safe0 = init_bd ? Num<T>::inf() : beta0;
        for (int i = tid; i < n; i += nt) V[i] = sr[i] / safe0;
        for (int idx = tid; idx < R * npad; idx += nt) V[npad + idx] = T(0);
        for (int idx = tid; idx < R * (R + 1); idx += nt)
          coeff[idx] = (idx / (R + 1) == idx % (R + 1)) ? T(1) : T(0);
        __syncthreads();
        bd_new = init_bd;
        for (int k = 0; k < R && !bd_new; ++k) {
          // _arnoldi_gram_schmidt, gmres.py:331-399
          if (Mg) {
            cta_matvec<T>(A, n, n, n, V + (size_t)k * npad, st, T(1));
            cta_matvec<T>(Mg, n, n, n, st, sw, T(1));
          } else {
            cta_matvec<T>(A, n, n, n, V + (size_t)k * npad, sw, T(1));
          }
          const T step_norm = cta_two_norm<T>(sw, n, red);
          // proj = V^H w over ALL restart+1 columns: one warp per column
          {
            const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
            for (int j = warp; j <= R; j += nw) {
              T a = T(0);
              const T* vj = V + (size_t)j * npad;
              for (int i = lane; i < n; i += 32) a = fma_(vj[i], sw[i], a);
              a = warp_sum(a);
              if (lane == 0) proj[j] = a;
            }
          }
          __syncthreads();
          for (int i = tid; i < n; i += nt) {
            T acc = T(0);
            for (int j = 0; j <= R; ++j) acc = fma_(V[(size_t)j * npad + i], proj[j], acc);
            sw[i] = sw[i] - acc;
          }
          __syncthreads();
          const T nrm = cta_two_norm<T>(sw, n, red);
          bd_new = nrm < step_norm * eps;
          const T safe = bd_new ? Num<T>::inf() : nrm;
          for (int i = tid; i < n; i += nt) V[(size_t)(k + 1) * npad + i] = sw[i] / safe;
          for (int j = tid; j <= R; j += nt) coeff[k * (R + 1) + j] = (j == k + 1) ? nrm : proj[j];
          __syncthreads();
        }
        // z = lstsq(coeff^T, [beta0, 0, ...]) (gmres.py:295-303)
        for (int idx = tid; idx < (R + 1) * R; idx += nt) {
          const int i = idx / R, c = idx % R;
          Qm[idx] = coeff[c * (R + 1) + i];
        }
        for (int i = tid; i <= R; i += nt) rhs[i] = i == 0 ? beta0 : T(0);
        __syncthreads();
        cta_hessenberg_lstsq<T>(Qm, rhs, zv, R, sc);
        for (int i = tid; i < n; i += nt) {
          T acc = T(0);
          for (int j = 0; j < R; ++j) acc = fma_(V[(size_t)j * npad + i], zv[j], acc);
          sd[i] = acc;
          sy[i] = sy[i] + acc;
        }
        diff_inf = false;
        __syncthreads();
      }
      // r = M (b - A y), gmres.py:318
      cta_matvec<T>(A, n, n, n, sy, st, T(1));
      if (Mg) {
        for (int i = tid; i < n; i += nt) st[i] = sb[i] - st[i];
        __syncthreads();
        cta_matvec<T>(Mg, n, n, n, st, sr, T(1));
      } else {
        for (int i = tid; i < n; i += nt) sr[i] = sb[i] - st[i];
        __syncthreads();
      }
      // stagnation bookkeeping, gmres.py:175-179 (norm = max_norm, NaN-propagating)
      T mx[1] = {T(0)};
      for (int i = tid; i < n; i += nt) mx[0] = absmax2(mx[0], sr[i]);
      block_absmax<T, 1>(mx, red);
      const T rn = mx[0];
      const bool decreased = (rn - r_min) < T(0);
      stag = decreased ? 0 : stag + 1;
      r_min = (rn < r_min || rn != rn) ? rn : r_min;  // jnp.minimum propagates NaN
      deferred = breakdown;
      breakdown = bd_new;
      step += 1;
    }
    int result = krylov_final_result(step, p.max_steps, p.flags, has_scale);
    if (stag >= p.stagnation_iters) result = LXB_STAGNATION;  // gmres.py:228-230
    const bool nc = cta_not_converged<T>(sr, sd, sy, sb, n, p.rtol, p.atol, has_scale, diff_inf, red + 32);
    if (deferred && nc) result = LXB_BREAKDOWN;  // gmres.py:235-237
    for (int i = tid; i < n; i += nt) p.x[sys * n + i] = sy[i];
    if (tid == 0) {
      p.result[sys] = result;
      p.num_steps[sys] = (int32_t)step;
    }
    __syncthreads();
  }
}

template <typename T>
struct GmresPlan {
  size_t fixed_bytes, basis_bytes, mat_bytes, smem;
  bool basis_smem, a_smem;
  int blocks_cap;
};

template <typename T>
GmresPlan<T> gmres_plan(int n, int R) {
  const size_t kMax = 227 * 1024;
  GmresPlan<T> pl{};
  const size_t npad = ((size_t)n + 3) & ~(size_t)3;
  pl.fixed_bytes = (6 * npad + 96 + (R + 1) + R + (R + 1) + 2 * (size_t)R * (R + 1) + 4) * sizeof(T) + 16;
  pl.basis_bytes = (size_t)(R + 1) * npad * sizeof(T);
  pl.mat_bytes = (size_t)n * n * sizeof(T);
  pl.basis_smem = pl.fixed_bytes + pl.basis_bytes <= kMax;
  pl.a_smem = pl.basis_smem && (pl.fixed_bytes + pl.basis_bytes + pl.mat_bytes <= kMax) && n > 0;
  pl.smem = pl.fixed_bytes + (pl.basis_smem ? pl.basis_bytes : 0) + (pl.a_smem ? pl.mat_bytes : 0);
  return pl;
}

template <typename T>
bool gmres_wants_grid(int64_t batch, int n, int restart) {
  if (restart > n) restart = n;
  if (restart + 2 > kGridMaxKHost) return false;
  return use_grid_tier(batch, n, n) || gmres_plan<T>(n, restart).fixed_bytes > 227 * 1024;
}

template <typename T>
int gmres_dispatch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (p.batch < 0 || p.n < 0 || p.restart < 0 || !p.A || !p.b || !p.x || !p.result || !p.num_steps)
    return LXB_E_BADARG;
  if (p.batch == 0) return 0;
  if (p.restart > p.n) p.restart = p.n;  // gmres.py:128
  if (p.n == 0) p.restart = 0;
  const GmresPlan<T> pl = gmres_plan<T>(p.n, p.restart);
  if (gmres_wants_grid<T>(p.batch, p.n, p.restart)) return gmres_grid_launch<T>(p, ws, ws_bytes, st);
  if (pl.fixed_bytes > 227 * 1024) return LXB_E_UNSUPPORTED;
  p.a_smem = pl.a_smem;
  auto kern = gmres_cta_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int occ = 1;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kKrylovThreads, pl.smem));
  if (occ < 1) occ = 1;
  int64_t cap = (int64_t)kNumSMs * occ;
  if (!p.a_smem) cap = l2_resident_cap(cap, pl.mat_bytes);
  const int64_t blocks = p.batch < cap ? p.batch : cap;
  if (pl.basis_smem) {
    p.ws = nullptr;
  } else {
    const size_t per = pl.basis_bytes / sizeof(T);
    if (!ws || ws_bytes < (size_t)blocks * pl.basis_bytes) return LXB_E_WORKSPACE;
    p.ws = reinterpret_cast<T*>(ws);
    p.ws_stride = (int64_t)per;
  }
  kern<<<(unsigned)blocks, kKrylovThreads, pl.smem, st>>>(p);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_GMRES(sfx, T)                                                                      \
  extern "C" int lxb_gmres_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b,       \
                                 const T* Minv, int64_t stride_M, T* x, int32_t* result,           \
                                 int32_t* num_steps, int64_t batch, int32_t n, T rtol, T atol,     \
                                 int32_t max_steps, int32_t restart, int32_t stagnation_iters,     \
                                 int32_t flags, void* workspace, size_t workspace_bytes,           \
                                 lxb_stream_t stream) {                                            \
    lxb::KrylovParams<T> p{};                                                                      \
    p.A = A; p.sA = stride_A; p.b = b; p.sb = stride_b; p.M = Minv; p.sM = stride_M; p.x = x;      \
    p.result = result; p.num_steps = num_steps; p.batch = batch; p.m = n; p.n = n;                 \
    p.rtol = rtol; p.atol = atol; p.max_steps = max_steps; p.restart = restart;                    \
    p.stagnation_iters = stagnation_iters; p.flags = flags;                                        \
    return lxb::gmres_dispatch<T>(p, workspace, workspace_bytes, (cudaStream_t)stream);            \
  }                                                                                                \
  extern "C" size_t lxb_gmres_workspace_##sfx(int64_t batch, int32_t n, int32_t restart) {         \
    if (batch <= 0 || n <= 0) return 0;                                                            \
    if (restart > n) restart = n;                                                                  \
    if (lxb::gmres_wants_grid<T>(batch, n, restart)) return lxb::gmres_grid_ws_bytes<T>(n, restart); \
    const lxb::GmresPlan<T> pl = lxb::gmres_plan<T>(n, restart);                                   \
    if (pl.basis_smem) return 0;                                                                   \
    const int64_t cap = (int64_t)lxb::kNumSMs * 8;                                                 \
    return (size_t)(batch < cap ? batch : cap) * pl.basis_bytes;                                   \
  }
LXB_DEF_GMRES(f32, float)
LXB_DEF_GMRES(f64, double)

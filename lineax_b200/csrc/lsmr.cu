// LSMR (damp = 0), lineax/_solver/lsmr.py:94-359 (a port of SciPy's lsmr), as one
// persistent kernel per batch: Golub-Kahan bidiagonalisation (A v and A^T u), the three
// Givens rotations, the h / hbar / x updates, the ||r||, ||A||, cond(A) estimates and the
// four stopping tests all run on chip.  Scalars are carried redundantly by every thread.
#include "krylov_cta.cuh"
#include "krylov_grid_api.cuh"

namespace lxb {

template <typename T>
__device__ __forceinline__ T sign_(T a) {
  return a > T(0) ? T(1) : (a < T(0) ? T(-1) : a);  // jnp.sign (0 -> 0, NaN -> NaN)
}

// lsmr.py:361-409
template <typename T>
__device__ __forceinline__ void givens(T a, T b, T& c, T& s, T& r) {
  if (a == T(0) || b == T(0)) {
    if (b == T(0)) {
      c = sign_(a); s = T(0); r = abs_(a);
    } else {
      c = T(0); s = sign_(b); r = abs_(b);
    }
  } else if (abs_(b) > abs_(a)) {
    const T tau = a / b;
    s = sign_(b) / sqrt_(T(1) + tau * tau);
    c = s * tau;
    r = b / (s == T(0) ? T(1) : s);
  } else {
    const T tau = b / a;
    c = sign_(a) / sqrt_(T(1) + tau * tau);
    s = c * tau;
    r = a / (c == T(0) ? T(1) : c);
  }
}

template <typename T>
__device__ __forceinline__ T cta_norm2(const T* x, int n, T* red) {
  if (n == 1) return abs_(x[0]);  // _norm.py:74-80
  T s[1] = {T(0)};
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[0] = fma_(x[i], x[i], s[0]);
  block_sum<T, 1>(s, red);
  return sqrt_(s[0]);
}

template <typename T>
__global__ void __launch_bounds__(kKrylovThreads) lsmr_cta_kernel(KrylovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = p.m, n = p.n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int mpad = (m + 3) & ~3, npad = (n + 3) & ~3;
  T* su = reinterpret_cast<T*>(smem_raw);
  T* st_ = su + mpad;  // A v scratch (m)
  T* sv = st_ + mpad;
  T* sx = sv + npad;
  T* sh = sx + npad;
  T* shb = sh + npad;
  T* sq = shb + npad;  // A^T u scratch (n)
  T* red = sq + npad;
  T* sA = red + 96;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));

  for (int64_t sys = blockIdx.x; sys < p.batch; sys += gridDim.x) {
    const T* A = p.A + sys * p.sA;
    if (p.a_smem) {
      cta_stage_matrix<T>(A, sA, (size_t)m * n);
      A = sA;
    }
    for (int i = tid; i < n; i += nt) {
      sx[i] = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + i] : T(0);
      shb[i] = T(0);
    }
    for (int i = tid; i < m; i += nt) su[i] = p.b[sys * p.sb + i];
    __syncthreads();
    const T normb = cta_norm2<T>(su, m, red);
    cta_matvec<T>(A, n, m, n, sx, st_, T(1));
    for (int i = tid; i < m; i += nt) su[i] = su[i] - st_[i];
    __syncthreads();
    T beta = cta_norm2<T>(su, m, red);
    T alpha = T(0);
    if (beta == T(0)) {  // lsmr.py:143-151
      for (int i = tid; i < n; i += nt) sv[i] = T(0);
      __syncthreads();
    } else {
      for (int i = tid; i < m; i += nt) su[i] = su[i] / beta;
      __syncthreads();
      cta_matvec_t<T>(A, n, m, n, su, sv, T(1));
      alpha = cta_norm2<T>(sv, n, red);
    }
    {
      const T den = alpha == T(0) ? T(1) : alpha;
      for (int i = tid; i < n; i += nt) {
        const T v = sv[i] / den;
        sv[i] = v;
        sh[i] = v;
      }
      __syncthreads();
    }
    int64_t itn = 0;
    T zetabar = alpha * beta, alphabar = alpha, rho = T(1), rhobar = T(1), cbar = T(1), sbar = T(0);
    T betadd = beta, betad = T(0), rhodold = T(1), tautildeold = T(0), thetatilde = T(0), zeta = T(0),
      delta = T(0);
    T normA2 = alpha * alpha, maxrbar = T(0), minrbar = Num<T>::max(), condA = T(1);
    int istop = 0;
    T normr = beta, normAr = alpha * beta;
    if (alpha == T(0)) istop = 2;
    if (beta == T(0)) istop = 1;

    while (istop == 0) {
      itn += 1;
      // bidiagonalisation, lsmr.py:214-237
      cta_matvec<T>(A, n, m, n, sv, st_, T(1));
      for (int i = tid; i < m; i += nt) su[i] = su[i] * -alpha + st_[i];
      __syncthreads();
      beta = cta_norm2<T>(su, m, red);
      if (beta != T(0)) {
        for (int i = tid; i < m; i += nt) su[i] = su[i] / beta;
        __syncthreads();
        cta_matvec_t<T>(A, n, m, n, su, sq, T(1));
        for (int i = tid; i < n; i += nt) sv[i] = sv[i] * -beta + sq[i];
        __syncthreads();
        alpha = cta_norm2<T>(sv, n, red);
        const T den = alpha == T(0) ? T(1) : alpha;
        for (int i = tid; i < n; i += nt) sv[i] = sv[i] / den;
        __syncthreads();
      }
      T chat, shat, alphahat;
      givens<T>(alphabar, T(0), chat, shat, alphahat);
      const T rhoold = rho;
      T c, s;
      givens<T>(alphahat, beta, c, s, rho);
      const T thetanew = s * alpha;
      alphabar = c * alpha;
      const T rhobarold = rhobar, zetaold = zeta;
      const T thetabar = sbar * rho;
      const T rhotemp = cbar * rho;
      givens<T>(cbar * rho, thetanew, cbar, sbar, rhobar);
      zeta = cbar * zetabar;
      zetabar = -sbar * zetabar;
      // h, hbar, x updates, lsmr.py:261-272
      const T f1 = -(thetabar * rho / (rhoold * rhobarold));
      const T f2 = zeta / (rho * rhobar);
      const T f3 = -(thetanew / rho);
      for (int i = tid; i < n; i += nt) {
        const T hb = shb[i] * f1 + sh[i];
        shb[i] = hb;
        sx[i] = sx[i] + f2 * hb;
        sh[i] = sh[i] * f3 + sv[i];
      }
      __syncthreads();
      // ||r|| estimate, lsmr.py:276-300
      const T betaacute = chat * betadd;
      const T betacheck = -shat * betadd;
      const T betahat = c * betaacute;
      betadd = -s * betaacute;
      const T thetatildeold = thetatilde;
      T ctildeold, stildeold, rhotildeold;
      givens<T>(rhodold, thetabar, ctildeold, stildeold, rhotildeold);
      thetatilde = stildeold * rhobar;  // lsmr.py:286 reads the UPDATED rhobar
      rhodold = ctildeold * rhobar;
      betad = -stildeold * betad + ctildeold * betahat;
      tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold;
      const T taud = (zeta - thetatilde * tautildeold) / rhodold;
      delta = delta + betacheck * betacheck;
      const T dd = betad - taud;
      normr = sqrt_(delta + dd * dd + betadd * betadd);
      // ||A||, cond(A), lsmr.py:303-314
      normA2 = normA2 + beta * beta;
      const T normA = sqrt_(normA2);
      normA2 = normA2 + alpha * alpha;
      maxrbar = (maxrbar > rhobarold || maxrbar != maxrbar) ? maxrbar : rhobarold;
      if (itn > 1) minrbar = (minrbar < rhobarold || minrbar != minrbar) ? minrbar : rhobarold;
      {
        const T mx = (maxrbar > rhotemp || maxrbar != maxrbar) ? maxrbar : rhotemp;
        const T mn = (minrbar < rhotemp || minrbar != minrbar) ? minrbar : rhotemp;
        condA = mx / mn;
      }
      normAr = abs_(zetabar);
      const T normx = cta_norm2<T>(sx, n, red);
      const T well_posed_tol = p.atol + p.rtol * (normA * normx + normb);
      const T least_squares_tol = p.atol + p.rtol * (normA * normr);
      if (itn >= p.max_steps) istop = 4;  // lsmr.py:317-329, overwrite order
      if (condA > p.conlim) istop = 3;
      if (normAr < least_squares_tol) istop = 2;
      if (normr < well_posed_tol) istop = 1;
    }
    const T normx_final = cta_norm2<T>(sx, n, red);
    int result = krylov_final_result(itn, p.max_steps, p.flags, has_scale);
    if (istop < 3) result = LXB_SUCCESSFUL;  // lsmr.py:356-357
    if (istop == 3) result = LXB_CONLIM;
    for (int i = tid; i < n; i += nt) p.x[sys * n + i] = sx[i];
    if (tid == 0) {
      p.result[sys] = result;
      p.num_steps[sys] = (int32_t)(itn > 2147483647 ? 2147483647 : itn);
      if (p.stats) {
        T* so = p.stats + sys * 8;
        so[0] = T(istop); so[1] = normr; so[2] = normAr; so[3] = sqrt_(normA2);
        so[4] = condA; so[5] = normx_final; so[6] = T(0); so[7] = T(0);
      }
    }
    __syncthreads();
  }
}

template <typename T>
int lsmr_dispatch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (p.batch < 0 || p.n < 0 || p.m < 0 || !p.A || !p.b || !p.x || !p.result || !p.num_steps)
    return LXB_E_BADARG;
  if (p.batch == 0) return 0;
  const size_t kMax = 227 * 1024;
  const size_t mpad = ((size_t)p.m + 3) & ~(size_t)3, npad = ((size_t)p.n + 3) & ~(size_t)3;
  const size_t vec_bytes = (2 * mpad + 5 * npad + 96) * sizeof(T);
  if (use_grid_tier(p.batch, p.m, p.n) || vec_bytes > kMax) return lsmr_grid_launch<T>(p, ws, ws_bytes, st);
  const size_t mat_bytes = (size_t)p.m * p.n * sizeof(T);
  p.a_smem = (vec_bytes + mat_bytes <= kMax) && mat_bytes > 0;
  const size_t smem = vec_bytes + (p.a_smem ? mat_bytes : 0);
  auto kern = lsmr_cta_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kKrylovThreads, smem));
  if (occ < 1) occ = 1;
  const int64_t cap = (int64_t)kNumSMs * occ;
  const int64_t blocks = p.batch < cap ? p.batch : cap;
  kern<<<(unsigned)blocks, kKrylovThreads, smem, st>>>(p);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_LSMR(sfx, T)                                                                       \
  extern "C" int lxb_lsmr_##sfx(const T* A, int64_t stride_A, const T* b, int64_t stride_b, T* x,  \
                                int32_t* result, int32_t* num_steps, T* stats, int64_t batch,      \
                                int32_t m, int32_t n, T rtol, T atol, T conlim, int64_t max_steps, \
                                int32_t flags, void* workspace, size_t workspace_bytes,            \
                                lxb_stream_t stream) {                                             \
    lxb::KrylovParams<T> p{};                                                                      \
    p.A = A; p.sA = stride_A; p.b = b; p.sb = stride_b; p.x = x; p.result = result;                \
    p.num_steps = num_steps; p.stats = stats; p.batch = batch; p.m = m; p.n = n; p.rtol = rtol;    \
    p.atol = atol; p.conlim = conlim; p.max_steps = max_steps; p.flags = flags;                    \
    return lxb::lsmr_dispatch<T>(p, workspace, workspace_bytes, (cudaStream_t)stream);                                         \
  }                                                                                                \
  extern "C" size_t lxb_lsmr_workspace_##sfx(int64_t batch, int32_t m, int32_t n) {                \
    const size_t vb = (2 * (((size_t)m + 3) & ~(size_t)3) + 5 * (((size_t)n + 3) & ~(size_t)3) + 96) * sizeof(T); \
    return (lxb::use_grid_tier(batch, m, n) || vb > 227 * 1024) ? lxb::lsmr_grid_ws_bytes<T>(m, n) : 0; \
  }
LXB_DEF_LSMR(f32, float)
LXB_DEF_LSMR(f64, double)

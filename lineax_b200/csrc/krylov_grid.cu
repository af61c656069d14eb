// Grid-cooperative tier (one persistent cooperative kernel, all SMs on one system at a time) of
// CG / BiCGStab / GMRES / LSMR.  Same algorithms, statement for statement, as the CTA tier
// (cg.cu, bicgstab.cu, gmres.cu, lsmr.cu) with the CTA-local loops restricted to this CTA's slice
// and the block reductions replaced by GridTeam all-reduces.  Used for single (or few) large
// systems: BASELINE configs[0] (CG 1024^2 f64), configs[3] (GMRES 32768^2 f32), configs[4] (LSMR
// 262144 x 4096 f32).
#include "krylov_grid.cuh"
#include "krylov_grid_api.cuh"
#include "lsmr_grid.cuh"

namespace lxb {

template <typename T>
__device__ __forceinline__ bool grid_not_converged(GridTeam<T>& team, const T* r, const T* diff,
                                                   const T* y, const T* b, int lo, int hi, T rtol,
                                                   T atol, bool has_scale, bool diff_inf) {
  if (!has_scale) {
    team.sync();
    return true;
  }
  T v[2] = {T(0), T(0)};
  for (int i = lo + team.tid; i < hi; i += team.nt) {
    const T bs = atol + rtol * abs_(b[i]);
    const T ys = atol + rtol * abs_(y[i]);
    const T d = diff_inf ? Num<T>::inf() : diff[i];
    v[0] = absmax2(v[0], r[i] / bs);
    v[1] = absmax2(v[1], d / ys);
  }
  team.template reduce<0, 2>(nullptr, v);
  return (v[0] > T(1)) || (v[1] > T(1));
}

// ------------------------------------------------------------------------ CG ----
template <typename T>
__global__ void __launch_bounds__(kGridThreads) cg_grid_kernel(KrylovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* red = reinterpret_cast<T*>(smem_raw);
  const int n = p.n;
  const size_t npad = ((size_t)n + 3) & ~(size_t)3;
  T* part = p.ws;
  GridTeam<T> team(part, red);
  T* wy = part + grid_part_elems();
  T* wr = wy + npad;
  T* wp = wr + npad;
  T* wq = wp + npad;
  T* wd = wq + npad;
  int lo, hi;
  team.slice(n, lo, hi);
  const int tid = team.tid, nt = team.nt;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const T sign = (p.flags & LXB_NSD) ? T(-1) : T(1);
  const T rcond = T(2) * Num<T>::eps() * T(n);

  for (int64_t sys = 0; sys < p.batch; ++sys) {
    const T* A = p.A + sys * p.sA;
    const T* Mg = p.M ? p.M + sys * p.sM : nullptr;
    const T* b = p.b + sys * p.sb;
    for (int i = lo + tid; i < hi; i += nt) wy[i] = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + i] : T(0);
    team.sync();
    grid_matvec<T>(A, n, lo, hi, wy, wq, sign);
    for (int i = lo + tid; i < hi; i += nt) wr[i] = b[i] - wq[i];  // own rows only: no barrier needed
    if (Mg) {
      team.sync();
      grid_matvec<T>(Mg, n, lo, hi, wr, wp, T(1));
    } else {
      for (int i = lo + tid; i < hi; i += nt) wp[i] = wr[i];
    }
    __syncthreads();
    T gamma = grid_dot<T>(team, wp, wr, lo, hi);  // barrier inside: p is now globally visible
    int64_t step = 0;
    bool diff_inf = true;
    while (true) {
      if (!(gamma > T(0))) break;
      if (!(step < p.max_steps)) break;
      if (!grid_not_converged<T>(team, wr, wd, wy, b, lo, hi, p.rtol, p.atol, has_scale, diff_inf)) break;
      grid_matvec<T>(A, n, lo, hi, wp, wq, sign);
      __syncthreads();
      const T ip = grid_dot<T>(team, wq, wp, lo, hi);
      T alpha = gamma / ip;
      if (!(abs_(ip) > T(100) * rcond * abs_(gamma))) alpha = Num<T>::nan();
      step += 1;
      const bool stable = p.stabilise_every == 1 ||
                          (p.stabilise_every > 1 && (step % p.stabilise_every) == 0);
      for (int i = lo + tid; i < hi; i += nt) {
        const T d = alpha * wp[i];
        wd[i] = d;
        wy[i] = wy[i] + d;
        if (!stable) wr[i] = wr[i] - alpha * wq[i];
      }
      diff_inf = false;
      if (stable) {
        team.sync();  // y complete everywhere
        grid_matvec<T>(A, n, lo, hi, wy, wq, sign);
        for (int i = lo + tid; i < hi; i += nt) wr[i] = b[i] - wq[i];
      }
      const T* z = wr;
      if (Mg) {
        team.sync();  // r complete everywhere
        grid_matvec<T>(Mg, n, lo, hi, wr, wq, T(1));
        z = wq;
      }
      __syncthreads();
      const T gn = grid_dot<T>(team, z, wr, lo, hi);
      const T beta = gn / gamma;
      gamma = gn;
      for (int i = lo + tid; i < hi; i += nt) wp[i] = z[i] + beta * wp[i];
      // p becomes globally visible at the barrier inside the next grid_not_converged()
    }
    for (int i = lo + tid; i < hi; i += nt) p.x[sys * n + i] = (p.flags & LXB_NSD) ? -wy[i] : wy[i];
    if (team.bid == 0 && tid == 0) {
      p.result[sys] = krylov_final_result(step, p.max_steps, p.flags, has_scale);
      p.num_steps[sys] = (int32_t)step;
    }
    team.sync();
  }
}

// ------------------------------------------------------------------ BiCGStab ----
template <typename T>
__device__ __forceinline__ bool bicg_breakdown_g(T omega, T alpha, T rho, bool x64) {
  if (x64) return omega == T(0) || alpha == T(0) || rho == T(0);
  const T t = T(1e-16);
  return omega < t || alpha < t || rho < t;
}

template <typename T>
__global__ void __launch_bounds__(kGridThreads) bicgstab_grid_kernel(KrylovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* red = reinterpret_cast<T*>(smem_raw);
  const int n = p.n;
  const size_t npad = ((size_t)n + 3) & ~(size_t)3;
  T* part = p.ws;
  GridTeam<T> team(part, red);
  T* wy = part + grid_part_elems();
  T* wr0 = wy + npad;
  T* wr = wr0 + npad;
  T* wp = wr + npad;
  T* wv = wp + npad;
  T* wss = wv + npad;
  T* wt = wss + npad;
  T* wx = wt + npad;
  T* wz = wx + npad;
  T* wd = wz + npad;
  int lo, hi;
  team.slice(n, lo, hi);
  const int tid = team.tid, nt = team.nt;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const bool x64 = (p.flags & LXB_X64_BREAKDOWN) != 0;

  for (int64_t sys = 0; sys < p.batch; ++sys) {
    const T* A = p.A + sys * p.sA;
    const T* Mg = p.M ? p.M + sys * p.sM : nullptr;
    const T* b = p.b + sys * p.sb;
    for (int i = lo + tid; i < hi; i += nt) {
      wy[i] = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + i] : T(0);
      wp[i] = T(0);
      wv[i] = T(0);
    }
    team.sync();
    grid_matvec<T>(A, n, lo, hi, wy, wt, T(1));
    for (int i = lo + tid; i < hi; i += nt) {
      const T r = b[i] - wt[i];
      wr0[i] = r;
      wr[i] = r;
    }
    __syncthreads();
    T alpha = T(1), omega = T(1), rho = T(1);
    int64_t step = 0;
    bool diff_inf = true;
    while (true) {
      if (bicg_breakdown_g(omega, alpha, rho, x64)) break;
      if (!grid_not_converged<T>(team, wr, wd, wy, b, lo, hi, p.rtol, p.atol, has_scale, diff_inf)) break;
      if (!(step < p.max_steps)) break;
      const T rho_new = grid_dot<T>(team, wr0, wr, lo, hi);
      const T beta = (rho_new / rho) * (alpha / omega);
      for (int i = lo + tid; i < hi; i += nt) wp[i] = wr[i] + beta * (wp[i] - omega * wv[i]);
      team.sync();  // p complete
      const T* xh = wp;
      if (Mg) {
        grid_matvec<T>(Mg, n, lo, hi, wp, wx, T(1));
        team.sync();
        xh = wx;
      }
      grid_matvec<T>(A, n, lo, hi, xh, wv, T(1));
      __syncthreads();
      alpha = rho_new / grid_dot<T>(team, wr0, wv, lo, hi);
      for (int i = lo + tid; i < hi; i += nt) wss[i] = wr[i] - alpha * wv[i];
      team.sync();  // s complete
      const T* z = wss;
      if (Mg) {
        grid_matvec<T>(Mg, n, lo, hi, wss, wz, T(1));
        team.sync();
        z = wz;
      }
      grid_matvec<T>(A, n, lo, hi, z, wt, T(1));
      __syncthreads();
      T d3[2] = {T(0), T(0)};
      for (int i = lo + tid; i < hi; i += nt) {
        d3[0] = fma_(wss[i], wt[i], d3[0]);
        d3[1] = fma_(wt[i], wt[i], d3[1]);
      }
      team.template reduce<2, 0>(d3, nullptr);
      omega = d3[0] / d3[1];
      for (int i = lo + tid; i < hi; i += nt) {
        const T d = alpha * xh[i] + omega * z[i];
        wd[i] = d;
        wy[i] = wy[i] + d;
        wr[i] = wss[i] - omega * wt[i];
      }
      diff_inf = false;
      rho = rho_new;
      step += 1;
      __syncthreads();
    }
    int result = krylov_final_result(step, p.max_steps, p.flags, has_scale);
    const bool nc = grid_not_converged<T>(team, wr, wd, wy, b, lo, hi, p.rtol, p.atol, has_scale, diff_inf);
    if (bicg_breakdown_g(omega, alpha, rho, x64) && nc) result = LXB_BREAKDOWN;
    for (int i = lo + tid; i < hi; i += nt) p.x[sys * n + i] = wy[i];
    if (team.bid == 0 && tid == 0) {
      p.result[sys] = result;
      p.num_steps[sys] = (int32_t)step;
    }
    team.sync();
  }
}

// --------------------------------------------------------------------- GMRES ----
// Every CTA keeps its own copy of the small Hessenberg data in shared memory and solves the
// least-squares problem redundantly (identical bits everywhere, no broadcast needed).
template <typename T>
__device__ void cta_hessenberg_lstsq_g(T* Q, T* rhs, T* z, int R, T* sc);

template <typename T>
__global__ void __launch_bounds__(kGridThreads) gmres_grid_kernel(KrylovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, R = p.restart;
  T* red = reinterpret_cast<T*>(smem_raw);  // 96 + kGridMaxK
  T* proj = red + 96 + kGridMaxK;           // R + 1 (<= kGridMaxK)
  T* zv = proj + kGridMaxK;
  T* rhs = zv + kGridMaxK;
  T* coeff = rhs + kGridMaxK;               // R x (R+1)
  T* Qm = coeff + (size_t)R * (R + 1);
  T* sc = Qm + (size_t)R * (R + 1);
  const size_t npad = ((size_t)n + 3) & ~(size_t)3;
  T* part = p.ws;
  GridTeam<T> team(part, red);
  T* wy = part + grid_part_elems();
  T* wr = wy + npad;
  T* ww = wr + npad;
  T* wd = ww + npad;
  T* wt = wd + npad;
  T* V = wt + npad;  // (R + 1) x npad
  int lo, hi;
  team.slice(n, lo, hi);
  const int tid = team.tid, nt = team.nt;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const T eps = Num<T>::eps();

  for (int64_t sys = 0; sys < p.batch; ++sys) {
    const T* A = p.A + sys * p.sA;
    const T* Mg = p.M ? p.M + sys * p.sM : nullptr;
    const T* b = p.b + sys * p.sb;
    for (int i = lo + tid; i < hi; i += nt) {
      wy[i] = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + i] : T(0);
      wr[i] = T(0);
    }
    __syncthreads();
    bool breakdown = false, deferred = false, diff_inf = true;
    T r_min = Num<T>::inf();
    int64_t step = 0;
    int stag = 0;
    while (true) {
      bool go = !deferred && stag < p.stagnation_iters;
      // evaluated unconditionally so that every CTA executes the same barriers
      const bool nc = grid_not_converged<T>(team, wr, wd, wy, b, lo, hi, p.rtol, p.atol, has_scale, diff_inf);
      go = (go && nc && step < p.max_steps) || step == 0;
      if (!go) break;
      bool bd_new = false;
      if (step > 0) {
        const T beta0 = grid_norm2<T>(team, wr, lo, hi, n);
        const bool init_bd = beta0 < eps;
        const T safe0 = init_bd ? Num<T>::inf() : beta0;
        for (int i = lo + tid; i < hi; i += nt) {
          V[i] = wr[i] / safe0;
          for (int j = 1; j <= R; ++j) V[(size_t)j * npad + i] = T(0);
        }
        for (int idx = tid; idx < R * (R + 1); idx += nt)
          coeff[idx] = (idx / (R + 1) == idx % (R + 1)) ? T(1) : T(0);
        team.sync();  // V[0] complete
        bd_new = init_bd;
        for (int k = 0; k < R && !bd_new; ++k) {
          if (Mg) {
            grid_matvec<T>(A, n, lo, hi, V + (size_t)k * npad, wt, T(1));
            team.sync();
            grid_matvec<T>(Mg, n, lo, hi, wt, ww, T(1));
          } else {
            grid_matvec<T>(A, n, lo, hi, V + (size_t)k * npad, ww, T(1));
          }
          __syncthreads();
          // local partials of ||w||^2 and of V^H w over all R+1 columns, one all-reduce for all
          {
            const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
            for (int j = warp; j <= R + 1; j += nw) {
              T a = T(0);
              const T* vj = j <= R ? V + (size_t)j * npad : ww;
              for (int i = lo + lane; i < hi; i += 32) a = fma_(vj[i], ww[i], a);
              a = warp_sum(a);
              if (lane == 0) proj[j] = a;  // proj[R+1] = ||w||^2 partial
            }
          }
          team.reduce_dyn(proj, R + 2);
          const T step_norm = n == 1 ? abs_(ww[0]) : sqrt_(proj[R + 1]);
          for (int i = lo + tid; i < hi; i += nt) {
            T acc = T(0);
            for (int j = 0; j <= R; ++j) acc = fma_(V[(size_t)j * npad + i], proj[j], acc);
            ww[i] = ww[i] - acc;
          }
          __syncthreads();
          const T nrm = grid_norm2<T>(team, ww, lo, hi, n);
          bd_new = nrm < step_norm * eps;
          const T safe = bd_new ? Num<T>::inf() : nrm;
          for (int i = lo + tid; i < hi; i += nt) V[(size_t)(k + 1) * npad + i] = ww[i] / safe;
          for (int j = tid; j <= R; j += nt) coeff[k * (R + 1) + j] = (j == k + 1) ? nrm : proj[j];
          team.sync();  // V[k+1] complete before the next matvec
        }
        for (int idx = tid; idx < (R + 1) * R; idx += nt) {
          const int i = idx / R, c = idx % R;
          Qm[idx] = coeff[c * (R + 1) + i];
        }
        for (int i = tid; i <= R; i += nt) rhs[i] = i == 0 ? beta0 : T(0);
        __syncthreads();
        cta_hessenberg_lstsq_g<T>(Qm, rhs, zv, R, sc);
        for (int i = lo + tid; i < hi; i += nt) {
          T acc = T(0);
          for (int j = 0; j < R; ++j) acc = fma_(V[(size_t)j * npad + i], zv[j], acc);
          wd[i] = acc;
          wy[i] = wy[i] + acc;
        }
        diff_inf = false;
      }
      team.sync();  // y complete
      grid_matvec<T>(A, n, lo, hi, wy, wt, T(1));
      if (Mg) {
        for (int i = lo + tid; i < hi; i += nt) wt[i] = b[i] - wt[i];
        team.sync();
        grid_matvec<T>(Mg, n, lo, hi, wt, wr, T(1));
      } else {
        for (int i = lo + tid; i < hi; i += nt) wr[i] = b[i] - wt[i];
      }
      __syncthreads();
      T mx[1] = {T(0)};
      for (int i = lo + tid; i < hi; i += nt) mx[0] = absmax2(mx[0], wr[i]);
      team.template reduce<0, 1>(nullptr, mx);
      const T rn = mx[0];
      const bool decreased = (rn - r_min) < T(0);
      stag = decreased ? 0 : stag + 1;
      r_min = (rn < r_min || rn != rn) ? rn : r_min;
      deferred = breakdown;
      breakdown = bd_new;
      step += 1;
    }
    int result = krylov_final_result(step, p.max_steps, p.flags, has_scale);
    if (stag >= p.stagnation_iters) result = LXB_STAGNATION;
    const bool nc = grid_not_converged<T>(team, wr, wd, wy, b, lo, hi, p.rtol, p.atol, has_scale, diff_inf);
    if (deferred && nc) result = LXB_BREAKDOWN;
    for (int i = lo + tid; i < hi; i += nt) p.x[sys * n + i] = wy[i];
    if (team.bid == 0 && tid == 0) {
      p.result[sys] = result;
      p.num_steps[sys] = (int32_t)step;
    }
    team.sync();
  }
}

// same routine as gmres.cu's cta_hessenberg_lstsq (kept local to this translation unit)
template <typename T>
__device__ void cta_hessenberg_lstsq_g(T* Q, T* rhs, T* z, int R, T* sc) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int rows = R + 1;
  for (int j = 0; j < R; ++j) {
    if (tid == 0) {
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
  if (tid == 0) {
    for (int k = R - 1; k >= 0; --k) {
      T s = rhs[k];
      for (int c = k + 1; c < R; ++c) s = fma_(-Q[k * R + c], z[c], s);
      z[k] = s / Q[k * R + k];
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------- LSMR ----
// (Givens rotation, fused Golub-Kahan pass and column reductions: lsmr_grid.cuh)

template <typename T>
__global__ void __launch_bounds__(kGridThreads) lsmr_grid_kernel(KrylovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* red = reinterpret_cast<T*>(smem_raw);
  const int m = p.m, n = p.n;
  const size_t mpad = ((size_t)m + 3) & ~(size_t)3, npad = ((size_t)n + 3) & ~(size_t)3;
  T* part = p.ws;
  GridTeam<T> team(part, red);
  T* wu = part + grid_part_elems();
  T* wt = wu + mpad;
  T* wv = wt + mpad;
  T* wx = wv + npad;
  T* wh = wx + npad;
  T* whb = wh + npad;
  T* pbuf = whb + npad;  // nb x npad partials of A^T u
  int rlo, rhi, clo, chi;
  team.slice(m, rlo, rhi);
  team.slice(n, clo, chi);
  const int tid = team.tid, nt = team.nt;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  constexpr int V = 16 / sizeof(T);
  const int ch = (n / V + nt - 1) / nt;  // 16-byte column chunks per thread

  for (int64_t sys = 0; sys < p.batch; ++sys) {
    const T* A = p.A + sys * p.sA;
    const T* b = p.b + sys * p.sb;
    // one-read-of-A path: rows must fit the per-thread register tile and be 16-byte aligned
    const bool fused = (n % V == 0) && ch >= 1 && ch <= 4 &&
                       ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && !(p.flags & (1 << 30));
    for (int i = clo + tid; i < chi; i += nt) {
      wx[i] = (p.flags & LXB_HAS_Y0) ? p.x[sys * n + i] : T(0);
      whb[i] = T(0);
      wv[i] = T(0);
    }
    for (int i = rlo + tid; i < rhi; i += nt) wu[i] = b[i];
    __syncthreads();
    const T normb = grid_norm2<T>(team, wu, rlo, rhi, m);  // barrier: x visible
    T beta, alpha = T(0);
    if (fused) {
      // u' = b - A x0 and A^T u' from one read of A
      T sq[1];
      sq[0] = lsmr_fused_dispatch<T>(ch, A, n, rlo, rhi, wx, wu, wu, T(-1), T(1),
                                     pbuf + (size_t)team.bid * npad, red);
      if (tid != 0) sq[0] = T(0);
      team.template reduce<1, 0>(sq, nullptr);  // barrier: partials visible
      beta = m == 1 ? abs_(wu[0]) : sqrt_(sq[0]);
      if (beta != T(0)) {
        for (int i = rlo + tid; i < rhi; i += nt) wu[i] = wu[i] / beta;
        grid_reduce_cols_scaled<T>(team, pbuf, (int)npad, clo, chi, wv, beta, T(0));
        __syncthreads();
        alpha = grid_norm2<T>(team, wv, clo, chi, n);
      }
    } else {
      grid_matvec<T>(A, n, rlo, rhi, wx, wt, T(1));
      for (int i = rlo + tid; i < rhi; i += nt) wu[i] = wu[i] - wt[i];
      __syncthreads();
      beta = grid_norm2<T>(team, wu, rlo, rhi, m);
      if (beta != T(0)) {
        for (int i = rlo + tid; i < rhi; i += nt) wu[i] = wu[i] / beta;
        __syncthreads();
        grid_matvec_t_partial<T>(A, n, rlo, rhi, wu, pbuf + (size_t)team.bid * npad);
        team.sync();
        grid_reduce_cols<T>(team, pbuf, (int)npad, clo, chi, wv, T(0));
        __syncthreads();
        alpha = grid_norm2<T>(team, wv, clo, chi, n);
      }
    }
    {
      const T den = alpha == T(0) ? T(1) : alpha;
      for (int i = clo + tid; i < chi; i += nt) {
        const T v = wv[i] / den;
        wv[i] = v;
        wh[i] = v;
      }
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
      team.sync();  // v complete
      if (fused) {
        T sq[1];
        sq[0] = lsmr_fused_dispatch<T>(ch, A, n, rlo, rhi, wv, wu, wu, T(1), -alpha,
                                       pbuf + (size_t)team.bid * npad, red);
        if (tid != 0) sq[0] = T(0);
        team.template reduce<1, 0>(sq, nullptr);
        beta = m == 1 ? abs_(wu[0]) : sqrt_(sq[0]);
        if (beta != T(0)) {
          for (int i = rlo + tid; i < rhi; i += nt) wu[i] = wu[i] / beta;
          grid_reduce_cols_scaled<T>(team, pbuf, (int)npad, clo, chi, wv, beta, -beta);
          __syncthreads();
          alpha = grid_norm2<T>(team, wv, clo, chi, n);
          const T den = alpha == T(0) ? T(1) : alpha;
          for (int i = clo + tid; i < chi; i += nt) wv[i] = wv[i] / den;
        }
      } else {
        grid_matvec<T>(A, n, rlo, rhi, wv, wt, T(1));
        for (int i = rlo + tid; i < rhi; i += nt) wu[i] = wu[i] * -alpha + wt[i];
        __syncthreads();
        beta = grid_norm2<T>(team, wu, rlo, rhi, m);
        if (beta != T(0)) {
          for (int i = rlo + tid; i < rhi; i += nt) wu[i] = wu[i] / beta;
          __syncthreads();
          grid_matvec_t_partial<T>(A, n, rlo, rhi, wu, pbuf + (size_t)team.bid * npad);
          team.sync();
          grid_reduce_cols<T>(team, pbuf, (int)npad, clo, chi, wv, -beta);
          __syncthreads();
          alpha = grid_norm2<T>(team, wv, clo, chi, n);
          const T den = alpha == T(0) ? T(1) : alpha;
          for (int i = clo + tid; i < chi; i += nt) wv[i] = wv[i] / den;
        }
      }
      T chat, shat, alphahat;
      givens_g<T>(alphabar, T(0), chat, shat, alphahat);
      const T rhoold = rho;
      T c, s;
      givens_g<T>(alphahat, beta, c, s, rho);
      const T thetanew = s * alpha;
      alphabar = c * alpha;
      const T rhobarold = rhobar, zetaold = zeta;
      const T thetabar = sbar * rho;
      const T rhotemp = cbar * rho;
      givens_g<T>(cbar * rho, thetanew, cbar, sbar, rhobar);
      zeta = cbar * zetabar;
      zetabar = -sbar * zetabar;
      const T f1 = -(thetabar * rho / (rhoold * rhobarold));
      const T f2 = zeta / (rho * rhobar);
      const T f3 = -(thetanew / rho);
      for (int i = clo + tid; i < chi; i += nt) {
        const T hb = whb[i] * f1 + wh[i];
        whb[i] = hb;
        wx[i] = wx[i] + f2 * hb;
        wh[i] = wh[i] * f3 + wv[i];
      }
      __syncthreads();
      const T betaacute = chat * betadd;
      const T betacheck = -shat * betadd;
      const T betahat = c * betaacute;
      betadd = -s * betaacute;
      const T thetatildeold = thetatilde;
      T ctildeold, stildeold, rhotildeold;
      givens_g<T>(rhodold, thetabar, ctildeold, stildeold, rhotildeold);
      thetatilde = stildeold * rhobar;
      rhodold = ctildeold * rhobar;
      betad = -stildeold * betad + ctildeold * betahat;
      tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold;
      const T taud = (zeta - thetatilde * tautildeold) / rhodold;
      delta = delta + betacheck * betacheck;
      const T dd = betad - taud;
      normr = sqrt_(delta + dd * dd + betadd * betadd);
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
      const T normx = grid_norm2<T>(team, wx, clo, chi, n);
      const T well_posed_tol = p.atol + p.rtol * (normA * normx + normb);
      const T least_squares_tol = p.atol + p.rtol * (normA * normr);
      if (itn >= p.max_steps) istop = 4;
      if (condA > p.conlim) istop = 3;
      if (normAr < least_squares_tol) istop = 2;
      if (normr < well_posed_tol) istop = 1;
    }
    const T normx_final = grid_norm2<T>(team, wx, clo, chi, n);
    int result = krylov_final_result(itn, p.max_steps, p.flags, has_scale);
    if (istop < 3) result = LXB_SUCCESSFUL;
    if (istop == 3) result = LXB_CONLIM;
    for (int i = clo + tid; i < chi; i += nt) p.x[sys * n + i] = wx[i];
    if (team.bid == 0 && tid == 0) {
      p.result[sys] = result;
      p.num_steps[sys] = (int32_t)(itn > 2147483647 ? 2147483647 : itn);
      if (p.stats) {
        T* so = p.stats + sys * 8;
        so[0] = T(istop); so[1] = normr; so[2] = normAr; so[3] = sqrt_(normA2);
        so[4] = condA; so[5] = normx_final; so[6] = T(0); so[7] = T(0);
      }
    }
    team.sync();
  }
}

// ------------------------------------------------------------------ launchers ----
// nb_max: the CTA count the workspace was sized for; the actual grid is the largest co-resident
// multiple of the SM count not exceeding it.
template <typename T, typename K>
int launch_grid(K kern, KrylovParams<T>& p, size_t smem, int nb_max, cudaStream_t st) {
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kGridThreads, smem));
  int dev = 0, sms = 0;
  LXB_CUDA_TRY(cudaGetDevice(&dev));
  LXB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (occ < 1) return LXB_E_UNSUPPORTED;
  int nb = occ * sms;  // all CTAs must be co-resident for the grid barrier
  if (nb > nb_max) nb = nb_max;
  void* args[] = {&p};
  LXB_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3(nb), dim3(kGridThreads), args, smem, st));
  count_launch();
  return 0;
}

template <typename T>
int cg_grid_launch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int nb = grid_blocks();
  if (!ws || ws_bytes < cg_grid_ws_bytes<T>(p.n)) return LXB_E_WORKSPACE;
  p.ws = reinterpret_cast<T*>(ws);
  return launch_grid<T>(cg_grid_kernel<T>, p, (96 + kGridMaxK) * sizeof(T), nb, st);
}
template <typename T>
int bicgstab_grid_launch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int nb = grid_blocks();
  if (!ws || ws_bytes < bicgstab_grid_ws_bytes<T>(p.n)) return LXB_E_WORKSPACE;
  p.ws = reinterpret_cast<T*>(ws);
  return launch_grid<T>(bicgstab_grid_kernel<T>, p, (96 + kGridMaxK) * sizeof(T), nb, st);
}
template <typename T>
int gmres_grid_launch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int nb = grid_blocks();
  if (p.restart + 2 > kGridMaxK) return LXB_E_UNSUPPORTED;
  if (!ws || ws_bytes < gmres_grid_ws_bytes<T>(p.n, p.restart)) return LXB_E_WORKSPACE;
  p.ws = reinterpret_cast<T*>(ws);
  const size_t smem = (96 + 4 * kGridMaxK + 2 * (size_t)p.restart * (p.restart + 1) + 8) * sizeof(T);
  return launch_grid<T>(gmres_grid_kernel<T>, p, smem, nb, st);
}
template <typename T>
int lsmr_grid_launch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int nb = grid_blocks();
  if (!ws || ws_bytes < lsmr_grid_ws_bytes<T>(p.m, p.n)) return LXB_E_WORKSPACE;
  p.ws = reinterpret_cast<T*>(ws);
  return launch_grid<T>(lsmr_grid_kernel<T>, p, (96 + kGridMaxK) * sizeof(T), nb, st);
}

#define LXB_INSTANTIATE_GRID(T)                                                                  \
  template int cg_grid_launch<T>(KrylovParams<T>, void*, size_t, cudaStream_t);                  \
  template int bicgstab_grid_launch<T>(KrylovParams<T>, void*, size_t, cudaStream_t);            \
  template int gmres_grid_launch<T>(KrylovParams<T>, void*, size_t, cudaStream_t);               \
  template int lsmr_grid_launch<T>(KrylovParams<T>, void*, size_t, cudaStream_t);
LXB_INSTANTIATE_GRID(float)
LXB_INSTANTIATE_GRID(double)

}  // namespace lxb

// Row-sharded CG across the GPUs of one NVLink/NVSwitch box ("also CG/BiCGStab"):
// rank p owns rows [row_offset, row_offset + n_local) of the SPD operator and the same slice of every
// vector.  ONE persistent cooperative kernel per GPU runs the whole solve (lineax/_solver/cg.py:114-227,
// no preconditioner); per iteration three exchange rounds of dist_team.cuh are fused into it over NVLink
// peer memory: all-gather of the search direction p (remote stores + round), all-reduce of <Ap, p>, and
// one round carrying <r, r> together with the two max-norms of the convergence test.  Every GPU folds
// the partials in rank order, so all ranks hold identical bits and the control flow cannot diverge.
#include "dist_team.cuh"

namespace lxb {

template <typename T>
struct CgDistParams {
  KrylovParams<T> k;
  unsigned char* const* peers;
  int world, rank, n_global, row_offset, stage_x;
};

template <typename T>
__global__ void __launch_bounds__(kDistThreads) cg_dist_kernel(CgDistParams<T> dp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const KrylovParams<T>& p = dp.k;
  const int n = dp.n_global, nl = p.n /* local rows */, off = dp.row_offset;
  T* red = reinterpret_cast<T*>(smem_raw);
  T* sc = red + 96 + kGridMaxK;  // 8 scalars
  T* xs = dp.stage_x ? reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(sc + 8) + 15) & ~(uintptr_t)15) : nullptr;
  const size_t lpad = ((size_t)nl + 3) & ~(size_t)3;
  T* part = p.ws;
  DistTeam<T> team(part, red, dp.peers, dp.world, dp.rank);
  GridTeam<T>& g = team.g;
  T* wy = part + grid_part_elems();
  T* wr = wy + lpad;
  T* wp = wr + lpad;
  T* wq = wp + lpad;
  T* wd = wq + lpad;
  int lo, hi;
  g.slice(nl, lo, hi);
  const int tid = g.tid, nt = g.nt;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const T sign = (p.flags & LXB_NSD) ? T(-1) : T(1);
  const T rcond = T(2) * Num<T>::eps() * T(n);
  const T* A = p.A;  // [nl, n] local rows
  const T* b = p.b;  // [nl]
  T* xfull = team.xchg();

  // all-gather of a row-distributed vector into the exchange buffer of every GPU, then q = sign * A x
  auto gather_matvec = [&](const T* v) {
    team.push(v, lo, hi, off);
    team.xround(sc, 0, sc, 0);
    grid_matvec<T>(A, n, lo, hi, xfull, wq, sign, xs);
  };
  // <a, c> over all GPUs
  auto xdot = [&](const T* a, const T* c) -> T {
    T v[1] = {T(0)};
    for (int i = lo + tid; i < hi; i += nt) v[0] = fma_(a[i], c[i], v[0]);
    block_sum<T, 1>(v, red);
    if (tid == 0) sc[0] = v[0];
    team.xround(sc, 1, sc, 0);
    return sc[0];
  };

  for (int i = lo + tid; i < hi; i += nt) wy[i] = (p.flags & LXB_HAS_Y0) ? p.x[i] : T(0);
  __syncthreads();
  gather_matvec(wy);  // r0 = b - A y0 (always evaluated, cg.py:128)
  for (int i = lo + tid; i < hi; i += nt) {
    wr[i] = b[i] - wq[i];
    wp[i] = wr[i];
  }
  __syncthreads();
  T gamma = xdot(wp, wr);
  // diff = +inf before the first step, so the convergence test cannot stop the loop yet (cg.py:149-167)
  T norm1 = Num<T>::inf(), norm2 = Num<T>::inf();
  int64_t step = 0;
  while (true) {
    if (!(gamma > T(0))) break;
    if (!(step < p.max_steps)) break;
    if (has_scale && !((norm1 > T(1)) || (norm2 > T(1)))) break;
    gather_matvec(wp);
    const T ip = xdot(wq, wp);
    T alpha = gamma / ip;
    if (!(abs_(ip) > T(100) * rcond * abs_(gamma))) alpha = Num<T>::nan();
    step += 1;
    const bool stable = p.stabilise_every == 1 || (p.stabilise_every > 1 && (step % p.stabilise_every) == 0);
    __syncthreads();
    for (int i = lo + tid; i < hi; i += nt) {
      const T d = alpha * wp[i];
      wd[i] = d;
      wy[i] = wy[i] + d;
      if (!stable) wr[i] = wr[i] - alpha * wq[i];
    }
    __syncthreads();
    if (stable) {  // cg.py:187-200: r = b - A y
      gather_matvec(wy);
      for (int i = lo + tid; i < hi; i += nt) wr[i] = b[i] - wq[i];
      __syncthreads();
    }
    // one round: <r, r> and the two max-norms of the convergence test
    {
      T s[1] = {T(0)}, v[2] = {T(0), T(0)};
      for (int i = lo + tid; i < hi; i += nt) {
        s[0] = fma_(wr[i], wr[i], s[0]);
        v[0] = absmax2(v[0], wr[i] / (p.atol + p.rtol * abs_(b[i])));
        v[1] = absmax2(v[1], wd[i] / (p.atol + p.rtol * abs_(wy[i])));
      }
      block_sum<T, 1>(s, red);
      block_absmax<T, 2>(v, red);
      if (tid == 0) { sc[0] = s[0]; sc[1] = v[0]; sc[2] = v[1]; }
      team.xround(sc, 1, sc + 1, 2);
    }
    const T gn = sc[0];
    norm1 = sc[1];
    norm2 = sc[2];
    __syncthreads();
    const T beta = gn / gamma;
    gamma = gn;
    for (int i = lo + tid; i < hi; i += nt) wp[i] = wr[i] + beta * wp[i];
    __syncthreads();
  }
  for (int i = lo + tid; i < hi; i += nt) p.x[i] = (p.flags & LXB_NSD) ? -wy[i] : wy[i];
  if (g.bid == 0 && tid == 0) {
    p.result[0] = krylov_final_result(step, p.max_steps, p.flags, has_scale);
    p.num_steps[0] = (int32_t)step;
  }
  team.finish(sc);
}

// Row-sharded BiCGStab (lineax/_solver/bicgstab.py:78-205, no preconditioner): same decomposition; per
// iteration two all-gathers (p and s, each fused with the exchange round that also makes it a barrier)
// and four all-reduce rounds (rho', <r0, v>, {<s, t>, <t, t>}, the two max-norms of the convergence test).
template <typename T>
__global__ void __launch_bounds__(kDistThreads) bicgstab_dist_kernel(CgDistParams<T> dp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const KrylovParams<T>& p = dp.k;
  const int n = dp.n_global, nl = p.n, off = dp.row_offset;
  T* red = reinterpret_cast<T*>(smem_raw);
  T* sc = red + 96 + kGridMaxK;
  T* xs = dp.stage_x ? reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(sc + 8) + 15) & ~(uintptr_t)15) : nullptr;
  const size_t lpad = ((size_t)nl + 3) & ~(size_t)3;
  T* part = p.ws;
  DistTeam<T> team(part, red, dp.peers, dp.world, dp.rank);
  GridTeam<T>& g = team.g;
  T* wy = part + grid_part_elems();
  T* wr0 = wy + lpad;
  T* wr = wr0 + lpad;
  T* wp = wr + lpad;
  T* wv = wp + lpad;
  T* wss = wv + lpad;
  T* wt = wss + lpad;
  T* wd = wt + lpad;
  int lo, hi;
  g.slice(nl, lo, hi);
  const int tid = g.tid, nt = g.nt;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const bool x64 = (p.flags & LXB_X64_BREAKDOWN) != 0;
  const T* A = p.A;
  const T* b = p.b;
  T* xfull = team.xchg();
  auto gather_matvec = [&](const T* v, T* out) {
    team.push(v, lo, hi, off);
    team.xround(sc, 0, sc, 0);
    grid_matvec<T>(A, n, lo, hi, xfull, out, T(1), xs);
  };
  auto xdot = [&](const T* a, const T* c) -> T {
    T v[1] = {T(0)};
    for (int i = lo + tid; i < hi; i += nt) v[0] = fma_(a[i], c[i], v[0]);
    block_sum<T, 1>(v, red);
    if (tid == 0) sc[0] = v[0];
    team.xround(sc, 1, sc, 0);
    const T r = sc[0];
    __syncthreads();
    return r;
  };
  auto breakdown = [&](T omega, T alpha, T rho) -> bool {
    if (x64) return omega == T(0) || alpha == T(0) || rho == T(0);
    const T t = T(1e-16);
    return omega < t || alpha < t || rho < t;  // signed test, bicgstab.py:110-113
  };
  // convergence test of bicgstab.py:115-126 over all GPUs
  auto not_converged = [&](bool diff_inf) -> bool {
    if (!has_scale) return true;
    T v[2] = {T(0), T(0)};
    for (int i = lo + tid; i < hi; i += nt) {
      const T d = diff_inf ? Num<T>::inf() : wd[i];
      v[0] = absmax2(v[0], wr[i] / (p.atol + p.rtol * abs_(b[i])));
      v[1] = absmax2(v[1], d / (p.atol + p.rtol * abs_(wy[i])));
    }
    block_absmax<T, 2>(v, red);
    if (tid == 0) { sc[1] = v[0]; sc[2] = v[1]; }
    team.xround(sc, 0, sc + 1, 2);
    const bool nc = (sc[1] > T(1)) || (sc[2] > T(1));
    __syncthreads();
    return nc;
  };

  for (int i = lo + tid; i < hi; i += nt) {
    wy[i] = (p.flags & LXB_HAS_Y0) ? p.x[i] : T(0);
    wp[i] = T(0);
    wv[i] = T(0);
  }
  __syncthreads();
  gather_matvec(wy, wt);
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
    if (breakdown(omega, alpha, rho)) break;
    if (!not_converged(diff_inf)) break;
    if (!(step < p.max_steps)) break;
    const T rho_new = xdot(wr0, wr);
    const T beta = (rho_new / rho) * (alpha / omega);
    for (int i = lo + tid; i < hi; i += nt) wp[i] = wr[i] + beta * (wp[i] - omega * wv[i]);
    __syncthreads();
    gather_matvec(wp, wv);
    alpha = rho_new / xdot(wr0, wv);
    for (int i = lo + tid; i < hi; i += nt) wss[i] = wr[i] - alpha * wv[i];
    __syncthreads();
    gather_matvec(wss, wt);
    {
      T d2[2] = {T(0), T(0)};
      for (int i = lo + tid; i < hi; i += nt) {
        d2[0] = fma_(wss[i], wt[i], d2[0]);
        d2[1] = fma_(wt[i], wt[i], d2[1]);
      }
      block_sum<T, 2>(d2, red);
      if (tid == 0) { sc[0] = d2[0]; sc[1] = d2[1]; }
      team.xround(sc, 2, sc, 0);
      omega = sc[0] / sc[1];
      __syncthreads();
    }
    for (int i = lo + tid; i < hi; i += nt) {
      const T d = alpha * wp[i] + omega * wss[i];
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
  const bool nc = not_converged(diff_inf);
  if (breakdown(omega, alpha, rho) && nc) result = LXB_BREAKDOWN;
  for (int i = lo + tid; i < hi; i += nt) p.x[i] = wy[i];
  if (g.bid == 0 && tid == 0) {
    p.result[0] = result;
    p.num_steps[0] = (int32_t)step;
  }
  team.finish(sc);
}

template <typename T>
size_t cg_dist_ws_bytes(int n_local) { return (grid_part_elems() + 8 * pad4(n_local)) * sizeof(T); }  // BiCGStab: 8 vectors

template <typename T>
int cg_dist_launch(const T* A_local, const T* b_local, T* x_local, int32_t* result, int32_t* num_steps, int n,
                   int n_local, int row_offset, T rtol, T atol, int max_steps, int stabilise_every, int flags,
                   void* ws, size_t ws_bytes, void* const* peers, int world, int rank, cudaStream_t st,
                   bool bicgstab = false) {
  if (!A_local || !b_local || !x_local || !result || !num_steps || !peers || n <= 0 || n_local < 0 ||
      world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return LXB_E_BADARG;
  if (!ws || ws_bytes < cg_dist_ws_bytes<T>(n_local)) return LXB_E_WORKSPACE;
  CgDistParams<T> dp{};
  dp.k.A = A_local; dp.k.b = b_local; dp.k.x = x_local; dp.k.result = result; dp.k.num_steps = num_steps;
  dp.k.batch = 1; dp.k.m = n_local; dp.k.n = n_local; dp.k.rtol = rtol; dp.k.atol = atol;
  dp.k.max_steps = max_steps; dp.k.stabilise_every = stabilise_every; dp.k.flags = flags;
  dp.k.ws = reinterpret_cast<T*>(ws);
  dp.peers = reinterpret_cast<unsigned char* const*>(peers);
  dp.world = world; dp.rank = rank; dp.n_global = n; dp.row_offset = row_offset;
  size_t smem = (96 + kGridMaxK + 8) * sizeof(T) + 16;
  const size_t xbytes = pad4(n) * sizeof(T);
  dp.stage_x = smem + xbytes <= 200 * 1024;
  if (dp.stage_x) smem += xbytes;
  auto kern = bicgstab ? bicgstab_dist_kernel<T> : cg_dist_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0, dev = 0, sms = 0;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDistThreads, smem));
  LXB_CUDA_TRY(cudaGetDevice(&dev));
  LXB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (occ < 1) return LXB_E_UNSUPPORTED;
  int nb = occ * sms;
  if (nb > grid_blocks()) nb = grid_blocks();
  void* args[] = {&dp};
  LXB_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3(nb), dim3(kDistThreads), args, smem, st));
  count_launch();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_CG_DIST(sfx, T)                                                                      \
  extern "C" int lxb_cg_rowsharded_##sfx(const T* A_local, const T* b_local, T* x_local,             \
                                         int32_t* result, int32_t* num_steps, int32_t n,             \
                                         int32_t n_local, int32_t row_offset, T rtol, T atol,        \
                                         int32_t max_steps, int32_t stabilise_every, int32_t flags,  \
                                         void* workspace, size_t workspace_bytes,                    \
                                         void* const* peer_buffers, int32_t world, int32_t rank,     \
                                         lxb_stream_t stream) {                                      \
    return lxb::cg_dist_launch<T>(A_local, b_local, x_local, result, num_steps, n, n_local,          \
                                  row_offset, rtol, atol, max_steps, stabilise_every, flags,         \
                                  workspace, workspace_bytes, peer_buffers, world, rank,             \
                                  (cudaStream_t)stream);                                             \
  }                                                                                                  \
  extern "C" size_t lxb_cg_rowsharded_workspace_##sfx(int32_t n_local) {                             \
    return lxb::cg_dist_ws_bytes<T>(n_local);                                                        \
  }                                                                                                  \
  extern "C" int lxb_bicgstab_rowsharded_##sfx(const T* A_local, const T* b_local, T* x_local,       \
                                               int32_t* result, int32_t* num_steps, int32_t n,       \
                                               int32_t n_local, int32_t row_offset, T rtol, T atol,  \
                                               int32_t max_steps, int32_t flags, void* workspace,    \
                                               size_t workspace_bytes, void* const* peer_buffers,    \
                                               int32_t world, int32_t rank, lxb_stream_t stream) {   \
    return lxb::cg_dist_launch<T>(A_local, b_local, x_local, result, num_steps, n, n_local,          \
                                  row_offset, rtol, atol, max_steps, 0, flags, workspace,            \
                                  workspace_bytes, peer_buffers, world, rank, (cudaStream_t)stream,  \
                                  true);                                                             \
  }
LXB_DEF_CG_DIST(f32, float)
LXB_DEF_CG_DIST(f64, double)

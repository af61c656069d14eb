// Row-sharded restarted GMRES across the GPUs of one NVLink/NVSwitch box (SURVEY.md section 8e,
// BASELINE configs[3]): rank p owns rows [row_offset, row_offset + n_local) of A and the same slice
// of every vector.  ONE persistent cooperative kernel per GPU runs the whole solve; the per-step
// exchanges are fused into it over NVLink peer memory (no NCCL call, no host round trip):
//   * all-gather of the next Krylov vector: every CTA PUSHES its slice into each peer's exchange
//     buffer with plain remote stores, then a cross-GPU barrier;
//   * all-reduce of the (restart + 2) Gram-Schmidt scalars / norms: each GPU pushes its partial
//     into a slot of every peer's buffer, barrier, every GPU sums the P slots in rank order
//     (deterministic, bit-identical on every rank, so control flow never diverges between GPUs).
// The cross-GPU barrier is a monotonically increasing epoch written with system-scope release
// stores into every peer's flag array and polled locally.
// Arithmetic and control flow are gmres.cu's (lineax/_solver/gmres.py:106-413).
#include "dist_team.cuh"

namespace lxb {

template <typename T>
struct DistParams {
  KrylovParams<T> k;
  unsigned char* const* peers;
  int world, rank, n_global, row_offset, stage_x;
};

template <typename T>
__device__ void cta_hessenberg_lstsq_d(T* Q, T* rhs, T* z, int R, T* sc) {
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

template <typename T>
__global__ void __launch_bounds__(kDistThreads) gmres_dist_kernel(DistParams<T> dp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const KrylovParams<T>& p = dp.k;
  const int n = dp.n_global, nl = p.n /* local rows */, R = p.restart, off = dp.row_offset;
  T* red = reinterpret_cast<T*>(smem_raw);
  T* proj = red + 96 + kGridMaxK;
  T* zv = proj + kGridMaxK;
  T* rhs = zv + kGridMaxK;
  T* coeff = rhs + kGridMaxK;
  T* Qm = coeff + (size_t)R * (R + 1);
  T* sc = Qm + (size_t)R * (R + 1);
  // optional staging area for the assembled matvec input (the symmetric exchange buffer is not
  // L1-cacheable): present when the launcher found room for n elements
  T* xs = dp.stage_x ? reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(sc + 8) + 15) & ~(uintptr_t)15) : nullptr;
  const size_t lpad = ((size_t)nl + 3) & ~(size_t)3;
  T* part = p.ws;
  DistTeam<T> team(part, red, dp.peers, dp.world, dp.rank);
  GridTeam<T>& g = team.g;
  T* wy = part + grid_part_elems();
  T* wr = wy + lpad;
  T* ww = wr + lpad;
  T* wd = ww + lpad;
  T* V = wd + lpad;  // (R + 1) x lpad, local slices
  int lo, hi;
  g.slice(nl, lo, hi);
  const int tid = g.tid, nt = g.nt;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  const T eps = Num<T>::eps();
  const T* A = p.A;  // [nl, n] local rows
  const T* b = p.b;  // [nl]
  T* xfull = team.xchg();

  // block-level partial of a dot over this CTA's slice -> smem slot
  auto cta_dot_to = [&](const T* a, const T* c, T* slot) {
    T v[1] = {T(0)};
    for (int i = lo + tid; i < hi; i += nt) v[0] = fma_(a[i], c[i], v[0]);
    block_sum<T, 1>(v, red);
    if (tid == 0) *slot = v[0];
  };
  // residual r = b - A y (y assembled in the exchange buffer) and the three max-norms the loop
  // needs: mx[0] = max|r| (stagnation), mx[1], mx[2] = the convergence test of gmres.py:130-141
  T* mx = zv;  // 3 slots of shared scratch (zv is only live inside the least-squares solve)
  auto residual_and_norms = [&](bool diff_inf) {
    team.push(wy, lo, hi, off);
    team.xround(sc, 0, sc, 0);  // y assembled everywhere
    grid_matvec<T>(A, n, lo, hi, xfull, ww, T(1), xs);
    T v[3] = {T(0), T(0), T(0)};
    for (int i = lo + tid; i < hi; i += nt) {
      const T r = b[i] - ww[i];
      wr[i] = r;
      const T bs = p.atol + p.rtol * abs_(b[i]);
      const T ys = p.atol + p.rtol * abs_(wy[i]);
      const T d = diff_inf ? Num<T>::inf() : wd[i];
      v[0] = absmax2(v[0], r);
      v[1] = absmax2(v[1], r / bs);
      v[2] = absmax2(v[2], d / ys);
    }
    block_absmax<T, 3>(v, red);
    if (tid == 0) { mx[0] = v[0]; mx[1] = v[1]; mx[2] = v[2]; }
    team.xround(sc, 0, mx, 3);
  };

  for (int i = lo + tid; i < hi; i += nt) {
    wy[i] = (p.flags & LXB_HAS_Y0) ? p.x[i] : T(0);
    wr[i] = T(0);
  }
  __syncthreads();
  bool breakdown = false, deferred = false, diff_inf = true, nc = true;
  T r_min = Num<T>::inf();
  int64_t step = 0;
  int stag = 0;
  while (true) {
    // cond_fun, gmres.py:143-158 (`nc` was evaluated on the current state at the end of the
    // previous pass; it is irrelevant at step 0)
    const bool go = (!deferred && stag < p.stagnation_iters && nc && step < p.max_steps) || step == 0;
    if (!go) break;
    bool bd_new = false;
    if (step > 0) {
      // V[0] = r / ||r||: the UNNORMALISED slice goes into the exchange buffer together with the
      // norm partial (one round); consumers scale the matvec by 1/||r|| instead
      team.push(wr, lo, hi, off);
      cta_dot_to(wr, wr, sc + 2);
      team.xround(sc + 2, 1, sc, 0);
      const T beta0 = sqrt_(sc[2]);
      const bool init_bd = beta0 < eps;
      T safe = init_bd ? Num<T>::inf() : beta0;
      for (int i = lo + tid; i < hi; i += nt) {
        V[i] = wr[i] / safe;
        for (int j = 1; j <= R; ++j) V[(size_t)j * lpad + i] = T(0);
      }
      for (int idx = tid; idx < R * (R + 1); idx += nt)
        coeff[idx] = (idx / (R + 1) == idx % (R + 1)) ? T(1) : T(0);
      __syncthreads();
      bd_new = init_bd;
      for (int k = 0; k < R && !bd_new; ++k) {
        grid_matvec<T>(A, n, lo, hi, xfull, ww, T(1) / safe, xs);
        {
          const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
          for (int j = warp; j <= R + 1; j += nw) {
            T a = T(0);
            const T* vj = j <= R ? V + (size_t)j * lpad : ww;
            for (int i = lo + lane; i < hi; i += 32) a = fma_(vj[i], ww[i], a);
            a = warp_sum(a);
            if (lane == 0) proj[j] = a;
          }
        }
        team.xround(proj, R + 2, sc, 0);
        const T step_norm = sqrt_(proj[R + 1]);
        for (int i = lo + tid; i < hi; i += nt) {
          T acc = T(0);
          for (int j = 0; j <= R; ++j) acc = fma_(V[(size_t)j * lpad + i], proj[j], acc);
          ww[i] = ww[i] - acc;
        }
        __syncthreads();
        team.push(ww, lo, hi, off);  // next Krylov vector, unnormalised
        cta_dot_to(ww, ww, sc + 2);
        team.xround(sc + 2, 1, sc, 0);
        const T nrm = sqrt_(sc[2]);
        bd_new = nrm < step_norm * eps;
        safe = bd_new ? Num<T>::inf() : nrm;
        for (int i = lo + tid; i < hi; i += nt) V[(size_t)(k + 1) * lpad + i] = ww[i] / safe;
        for (int j = tid; j <= R; j += nt) coeff[k * (R + 1) + j] = (j == k + 1) ? nrm : proj[j];
        __syncthreads();
      }
      for (int idx = tid; idx < (R + 1) * R; idx += nt) {
        const int i = idx / R, c = idx % R;
        Qm[idx] = coeff[c * (R + 1) + i];
      }
      for (int i = tid; i <= R; i += nt) rhs[i] = i == 0 ? beta0 : T(0);
      __syncthreads();
      cta_hessenberg_lstsq_d<T>(Qm, rhs, zv, R, sc);
      for (int i = lo + tid; i < hi; i += nt) {
        T acc = T(0);
        for (int j = 0; j < R; ++j) acc = fma_(V[(size_t)j * lpad + i], zv[j], acc);
        wd[i] = acc;
        wy[i] = wy[i] + acc;
      }
      diff_inf = false;
      __syncthreads();
    }
    residual_and_norms(diff_inf);
    const T rn = mx[0];
    nc = !has_scale || (mx[1] > T(1)) || (mx[2] > T(1));
    __syncthreads();
    const bool decreased = (rn - r_min) < T(0);
    stag = decreased ? 0 : stag + 1;
    r_min = (rn < r_min || rn != rn) ? rn : r_min;
    deferred = breakdown;
    breakdown = bd_new;
    step += 1;
  }
  int result = krylov_final_result(step, p.max_steps, p.flags, has_scale);
  if (stag >= p.stagnation_iters) result = LXB_STAGNATION;
  if (deferred && nc) result = LXB_BREAKDOWN;
  for (int i = lo + tid; i < hi; i += nt) p.x[i] = wy[i];
  if (g.bid == 0 && tid == 0) {
    p.result[0] = result;
    p.num_steps[0] = (int32_t)step;
  }
  team.finish(sc);
}

template <typename T>
size_t gmres_dist_ws_bytes(int n_local, int restart) {
  return (grid_part_elems() + (4 + (size_t)restart + 1) * pad4(n_local)) * sizeof(T);
}

template <typename T>
int gmres_dist_launch(const T* A_local, const T* b_local, T* x_local, int32_t* result,
                      int32_t* num_steps, int n, int n_local, int row_offset, T rtol, T atol,
                      int max_steps, int restart, int stagnation_iters, int flags, void* ws,
                      size_t ws_bytes, void* const* peers, int world, int rank, cudaStream_t st) {
  if (!A_local || !b_local || !x_local || !result || !num_steps || !peers || n <= 0 || n_local < 0 ||
      world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return LXB_E_BADARG;
  if (restart > n) restart = n;
  if (restart + 2 > kGridMaxKHost) return LXB_E_UNSUPPORTED;
  if (!ws || ws_bytes < gmres_dist_ws_bytes<T>(n_local, restart)) return LXB_E_WORKSPACE;
  DistParams<T> dp{};
  dp.k.A = A_local; dp.k.b = b_local; dp.k.x = x_local; dp.k.result = result; dp.k.num_steps = num_steps;
  dp.k.batch = 1; dp.k.m = n_local; dp.k.n = n_local; dp.k.rtol = rtol; dp.k.atol = atol;
  dp.k.max_steps = max_steps; dp.k.restart = restart; dp.k.stagnation_iters = stagnation_iters;
  dp.k.flags = flags; dp.k.ws = reinterpret_cast<T*>(ws);
  dp.peers = reinterpret_cast<unsigned char* const*>(peers);
  dp.world = world; dp.rank = rank; dp.n_global = n; dp.row_offset = row_offset;
  size_t smem = (96 + 4 * kGridMaxK + 2 * (size_t)restart * (restart + 1) + 8) * sizeof(T) + 16;
  const size_t xbytes = pad4(n) * sizeof(T);
  dp.stage_x = smem + xbytes <= 200 * 1024;
  if (dp.stage_x) smem += xbytes;
  auto kern = gmres_dist_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0, dev = 0, sms = 0;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDistThreads, smem));
  LXB_CUDA_TRY(cudaGetDevice(&dev));
  LXB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (occ < 1) return LXB_E_UNSUPPORTED;
  int nb = occ * sms;
  if (nb > grid_blocks()) nb = grid_blocks();
  if (const char* e = getenv("LXB_MV_RB")) {
    const int rb = atoi(e);
    LXB_CUDA_TRY(cudaMemcpyToSymbolAsync(g_mv_force_rb, &rb, sizeof(int), 0, cudaMemcpyHostToDevice, st));
  }
  void* args[] = {&dp};
  LXB_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3(nb), dim3(kDistThreads), args, smem, st));
  count_launch();
  return 0;
}

}  // namespace lxb

#define LXB_DEF_GMRES_DIST(sfx, T)                                                                  \
  extern "C" int lxb_gmres_rowsharded_##sfx(                                                         \
      const T* A_local, const T* b_local, T* x_local, int32_t* result, int32_t* num_steps, int32_t n, \
      int32_t n_local, int32_t row_offset, T rtol, T atol, int32_t max_steps, int32_t restart,        \
      int32_t stagnation_iters, int32_t flags, void* workspace, size_t workspace_bytes,               \
      void* const* peer_buffers, int32_t world, int32_t rank, lxb_stream_t stream) {                 \
    return lxb::gmres_dist_launch<T>(A_local, b_local, x_local, result, num_steps, n, n_local,       \
                                     row_offset, rtol, atol, max_steps, restart, stagnation_iters,   \
                                     flags, workspace, workspace_bytes, peer_buffers, world, rank,   \
                                     (cudaStream_t)stream);                                          \
  }                                                                                                  \
  extern "C" size_t lxb_gmres_rowsharded_workspace_##sfx(int32_t n_local, int32_t restart) {         \
    return lxb::gmres_dist_ws_bytes<T>(n_local, restart);                                            \
  }                                                                                                  \
  extern "C" size_t lxb_gmres_rowsharded_symm_bytes_##sfx(int32_t n) { return lxb::symm_bytes<T>(n); }
LXB_DEF_GMRES_DIST(f32, float)
LXB_DEF_GMRES_DIST(f64, double)

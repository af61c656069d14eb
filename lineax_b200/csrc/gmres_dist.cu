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
#include "krylov_grid.cuh"
#include "krylov_grid_api.cuh"

namespace lxb {

constexpr int kMaxPeers = 16;
constexpr size_t kSymmFlagBytes = 4096;  // [0]: persistent epoch, [64 + 16*r]: arrival slot of rank r

__host__ __device__ inline size_t symm_part_off() { return kSymmFlagBytes; }
template <typename T>
__host__ __device__ inline size_t symm_xchg_off() {
  return kSymmFlagBytes + (((size_t)2 * kGridMaxKHost * kMaxPeers * sizeof(T)) + 255) / 256 * 256;
}
template <typename T>
size_t symm_bytes(int n) { return symm_xchg_off<T>() + pad4(n) * sizeof(T) + 256; }

template <typename T>
struct DistTeam {
  GridTeam<T> g;
  int P, rank;
  unsigned char* const* peers;  // device array of P symmetric-buffer base pointers
  unsigned char* mine;
  unsigned long long epoch;
  int xflip;

  __device__ DistTeam(T* part, T* red, unsigned char* const* peers_, int P_, int rank_)
      : g(part, red), P(P_), rank(rank_), peers(peers_), mine(peers_[rank_]), xflip(0) {
    epoch = *reinterpret_cast<volatile unsigned long long*>(mine);
  }

  // Cross-GPU barrier. All remote stores issued by any thread of this GPU before the call are
  // visible to every peer after it (threads fence at system scope before the grid barrier).
  __device__ void xsync() {
    __threadfence_system();
    g.sync();
    epoch += 1;
    if (g.bid == 0 && g.tid < P) {
      unsigned long long* slot = reinterpret_cast<unsigned long long*>(peers[g.tid] + 64 + 16 * rank);
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(epoch) : "memory");
      const unsigned long long* my = reinterpret_cast<const unsigned long long*>(mine + 64 + 16 * g.tid);
      unsigned long long seen;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(my) : "memory");
      } while (seen < epoch);
    }
    g.sync();
  }

  // push my slice [lo, hi) (local indices) of a vector into every GPU's exchange buffer at the
  // global position; followed by xsync() the full vector is readable locally at xchg().
  __device__ T* xchg() const { return reinterpret_cast<T*>(mine + symm_xchg_off<T>()); }
  __device__ void push(const T* local, int lo, int hi, int row_offset) {
    for (int q = 0; q < P; ++q) {
      T* dst = reinterpret_cast<T*>(peers[q] + symm_xchg_off<T>()) + row_offset;
      for (int i = lo + g.tid; i < hi; i += g.nt) dst[i] = local[i];
    }
  }

  // global all-reduce of K sums held in shared memory `vals` (already summed over this GPU)
  __device__ void xreduce_sum(T* vals, int K) {
    T* slotbase = nullptr;
    const size_t off = symm_part_off() + (size_t)xflip * kGridMaxK * kMaxPeers * sizeof(T);
    xflip ^= 1;
    if (g.bid == 0) {
      for (int idx = g.tid; idx < K * P; idx += g.nt) {
        const int q = idx / K, k = idx % K;
        slotbase = reinterpret_cast<T*>(peers[q] + off);
        slotbase[(size_t)k * kMaxPeers + rank] = vals[k];
      }
    }
    xsync();
    const T* in = reinterpret_cast<const T*>(mine + off);
    __syncthreads();
    for (int k = g.tid; k < K; k += g.nt) {
      T s = T(0);
      for (int q = 0; q < P; ++q) s += __ldcg(in + (size_t)k * kMaxPeers + q);
      vals[k] = s;
    }
    __syncthreads();
  }
  __device__ T xreduce_max1(T v) {  // NaN-propagating abs-max of one value per GPU
    const size_t off = symm_part_off() + (size_t)xflip * kGridMaxK * kMaxPeers * sizeof(T);
    xflip ^= 1;
    if (g.bid == 0 && g.tid < P) reinterpret_cast<T*>(peers[g.tid] + off)[rank] = v;
    xsync();
    const T* in = reinterpret_cast<const T*>(mine + off);
    T m = T(0);
    for (int q = 0; q < P; ++q) m = absmax2(m, __ldcg(in + q));
    return m;
  }
  __device__ void finish() {
    xsync();  // nobody leaves (and starts overwriting exchange slots) while a peer still reads
    if (g.bid == 0 && g.tid == 0) *reinterpret_cast<volatile unsigned long long*>(mine) = epoch;
  }
};

template <typename T>
struct DistParams {
  KrylovParams<T> k;
  unsigned char* const* peers;
  int world, rank, n_global, row_offset;
};

// local + global dot over the distributed vector slices
template <typename T>
__device__ __forceinline__ void dist_sums(DistTeam<T>& t, T* vals_smem, int K) {
  t.g.reduce_dyn(vals_smem, K);  // over this GPU's CTAs
  t.xreduce_sum(vals_smem, K);   // over GPUs
}

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
__global__ void __launch_bounds__(kGridThreads) gmres_dist_kernel(DistParams<T> dp) {
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

  auto not_converged = [&](bool diff_inf) -> bool {
    if (!has_scale) {
      team.xsync();
      return true;
    }
    T v[2] = {T(0), T(0)};
    for (int i = lo + tid; i < hi; i += nt) {
      const T bs = p.atol + p.rtol * abs_(b[i]);
      const T ys = p.atol + p.rtol * abs_(wy[i]);
      const T d = diff_inf ? Num<T>::inf() : wd[i];
      v[0] = absmax2(v[0], wr[i] / bs);
      v[1] = absmax2(v[1], d / ys);
    }
    g.template reduce<0, 2>(nullptr, v);
    const T m0 = team.xreduce_max1(v[0]);
    const T m1 = team.xreduce_max1(v[1]);
    return (m0 > T(1)) || (m1 > T(1));
  };
  // global two-norm of a distributed vector (size-1 shortcut on the GLOBAL size)
  auto norm2 = [&](const T* a) -> T {
    T s[1] = {T(0)};
    for (int i = lo + tid; i < hi; i += nt) s[0] = fma_(a[i], a[i], s[0]);
    g.template reduce<1, 0>(s, nullptr);
    if (tid == 0) sc[2] = s[0];
    __syncthreads();
    team.xreduce_sum(sc + 2, 1);
    return sqrt_(sc[2]);  // (n == 1: sqrt(x^2) = |x|, the _norm.py:74-80 shortcut)
  };

  for (int i = lo + tid; i < hi; i += nt) {
    wy[i] = (p.flags & LXB_HAS_Y0) ? p.x[i] : T(0);
    wr[i] = T(0);
  }
  __syncthreads();
  bool breakdown = false, deferred = false, diff_inf = true;
  T r_min = Num<T>::inf();
  int64_t step = 0;
  int stag = 0;
  while (true) {
    bool go = !deferred && stag < p.stagnation_iters;
    const bool nc = not_converged(diff_inf);
    go = (go && nc && step < p.max_steps) || step == 0;
    if (!go) break;
    bool bd_new = false;
    if (step > 0) {
      const T beta0 = norm2(wr);
      const bool init_bd = beta0 < eps;
      const T safe0 = init_bd ? Num<T>::inf() : beta0;
      for (int i = lo + tid; i < hi; i += nt) {
        V[i] = wr[i] / safe0;
        for (int j = 1; j <= R; ++j) V[(size_t)j * lpad + i] = T(0);
      }
      for (int idx = tid; idx < R * (R + 1); idx += nt)
        coeff[idx] = (idx / (R + 1) == idx % (R + 1)) ? T(1) : T(0);
      __syncthreads();
      team.push(V, lo, hi, off);
      team.xsync();  // V[0] assembled in every GPU's exchange buffer
      bd_new = init_bd;
      for (int k = 0; k < R && !bd_new; ++k) {
        grid_matvec<T>(A, n, lo, hi, xfull, ww, T(1));
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
        dist_sums<T>(team, proj, R + 2);
        const T step_norm = sqrt_(proj[R + 1]);
        for (int i = lo + tid; i < hi; i += nt) {
          T acc = T(0);
          for (int j = 0; j <= R; ++j) acc = fma_(V[(size_t)j * lpad + i], proj[j], acc);
          ww[i] = ww[i] - acc;
        }
        __syncthreads();
        const T nrm = norm2(ww);
        bd_new = nrm < step_norm * eps;
        const T safe = bd_new ? Num<T>::inf() : nrm;
        for (int i = lo + tid; i < hi; i += nt) V[(size_t)(k + 1) * lpad + i] = ww[i] / safe;
        for (int j = tid; j <= R; j += nt) coeff[k * (R + 1) + j] = (j == k + 1) ? nrm : proj[j];
        __syncthreads();
        team.push(V + (size_t)(k + 1) * lpad, lo, hi, off);
        team.xsync();
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
    team.push(wy, lo, hi, off);
    team.xsync();  // y assembled everywhere
    grid_matvec<T>(A, n, lo, hi, xfull, ww, T(1));
    for (int i = lo + tid; i < hi; i += nt) wr[i] = b[i] - ww[i];
    __syncthreads();
    T mx[1] = {T(0)};
    for (int i = lo + tid; i < hi; i += nt) mx[0] = absmax2(mx[0], wr[i]);
    g.template reduce<0, 1>(nullptr, mx);
    const T rn = team.xreduce_max1(mx[0]);
    const bool decreased = (rn - r_min) < T(0);
    stag = decreased ? 0 : stag + 1;
    r_min = (rn < r_min || rn != rn) ? rn : r_min;
    deferred = breakdown;
    breakdown = bd_new;
    step += 1;
  }
  int result = krylov_final_result(step, p.max_steps, p.flags, has_scale);
  if (stag >= p.stagnation_iters) result = LXB_STAGNATION;
  const bool nc = not_converged(diff_inf);
  if (deferred && nc) result = LXB_BREAKDOWN;
  for (int i = lo + tid; i < hi; i += nt) p.x[i] = wy[i];
  if (g.bid == 0 && tid == 0) {
    p.result[0] = result;
    p.num_steps[0] = (int32_t)step;
  }
  team.finish();
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
  const size_t smem = (96 + 4 * kGridMaxK + 2 * (size_t)restart * (restart + 1) + 8) * sizeof(T);
  auto kern = gmres_dist_kernel<T>;
  LXB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0, dev = 0, sms = 0;
  LXB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kGridThreads, smem));
  LXB_CUDA_TRY(cudaGetDevice(&dev));
  LXB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (occ < 1) return LXB_E_UNSUPPORTED;
  int nb = occ * sms;
  if (nb > grid_blocks()) nb = grid_blocks();
  void* args[] = {&dp};
  LXB_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3(nb), dim3(kGridThreads), args, smem, st));
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

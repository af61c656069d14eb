// Row-sharded LSMR across the GPUs of one NVLink/NVSwitch box ("tall least
// squares": BASELINE configs[4] at 2/4/8 GPUs).  Rank p owns a contiguous block of ROWS of A and
// the same slice of b and u; the length-n vectors v, x, h, hbar are replicated and every GPU runs
// the identical scalar recurrence, so control flow never diverges between GPUs.
//
// ONE persistent cooperative kernel per GPU runs the whole solve.  Per iteration each GPU reads
// its rows of A exactly once (the fused Golub-Kahan pass of lsmr_grid.cuh: u' = A v - alpha u and
// the partial A^T u' together) and there is ONE cross-GPU exchange, fused into the kernel over
// peer memory: every GPU pushes its n partial column sums into a slot of each peer's exchange
// buffer, the round's all-reduce carries sum(u'^2), and after its barrier every GPU adds the P
// slots in rank order (bit-identical everywhere).  No NCCL call, no host round trip.
// Arithmetic and control flow are lsmr.cu's / krylov_grid.cu's (lineax/_solver/lsmr.py:94-409).
#include "dist_team.cuh"
#include "lsmr_grid.cuh"

namespace lxb {

template <typename T>
struct LsmrDistParams {
  KrylovParams<T> k;  // m = local rows, n = columns, A/b = local blocks, x = full solution (replicated)
  unsigned char* const* peers;
  int world, rank, m_global;
};

template <typename T>
size_t lsmr_symm_bytes(int n, int world) {
  return symm_xchg_off<T>() + (size_t)2 * world * pad4(n) * sizeof(T) + 256;
}

template <typename T>
__global__ void __launch_bounds__(kDistThreads) lsmr_dist_kernel(LsmrDistParams<T> dp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const KrylovParams<T>& p = dp.k;
  const int ml = p.m, n = p.n, P = dp.world;
  T* red = reinterpret_cast<T*>(smem_raw);
  T* sc = red + 96 + kGridMaxK;  // 8 scalars of shared scratch
  const size_t mpad = ((size_t)ml + 3) & ~(size_t)3, npad = ((size_t)n + 3) & ~(size_t)3;
  T* part = p.ws;
  DistTeam<T> team(part, red, dp.peers, dp.world, dp.rank);
  GridTeam<T>& g = team.g;
  T* wu = part + grid_part_elems();
  T* wv = wu + mpad;
  T* wx = wv + npad;
  T* wh = wx + npad;
  T* whb = wh + npad;
  T* wl = whb + npad;   // this GPU's column sums of A^T u'
  T* pbuf = wl + npad;  // nb x npad per-CTA partials
  int rlo, rhi, clo, chi;
  g.slice(ml, rlo, rhi);
  g.slice(n, clo, chi);
  const int tid = g.tid, nt = g.nt;
  const bool has_scale = !(p.rtol == T(0) && p.atol == T(0));
  constexpr int V = 16 / sizeof(T);
  const int ch = (n / V + nt - 1) / nt;  // 16-byte column chunks per thread (launcher: 1..4)
  const T* A = p.A;
  const T* b = p.b;
  int par = 0;  // parity of the exchange slots (a fast GPU may be one round ahead of a slow reader)

  // sum over ALL GPUs of the squares of a row-distributed vector
  auto xsumsq_rows = [&](const T* a) -> T {
    T v[1] = {T(0)};
    for (int i = rlo + tid; i < rhi; i += nt) v[0] = fma_(a[i], a[i], v[0]);
    block_sum<T, 1>(v, red);
    if (tid == 0) sc[0] = v[0];
    team.xround(sc, 1, sc, 0);
    return sc[0];
  };
  // After a fused pass (per-CTA partials in pbuf, `ssq` = this CTA's sum of u'^2):
  // beta = ||u'|| over all GPUs; u' /= beta; v = (A^T u') / beta - beta * v  (v = 0 on the first call)
  auto exchange = [&](T ssq, bool first, T& beta) {
    if (tid == 0) sc[0] = ssq;
    g.sync();  // every CTA's partial column sums are in pbuf
    {
      const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
      for (int j = clo + warp; j < chi; j += nw) {
        T acc = T(0);
        for (int c = lane; c < g.nb; c += 32) acc += __ldcg(pbuf + (size_t)c * npad + j);
        acc = warp_sum(acc);
        if (lane == 0) wl[j] = acc;
      }
    }
    __syncthreads();
    for (int q = 0; q < P; ++q) {
      T* dst = reinterpret_cast<T*>(team.peers[q] + symm_xchg_off<T>()) + ((size_t)par * P + team.rank) * npad;
      for (int j = clo + tid; j < chi; j += nt) dst[j] = wl[j];
    }
    team.xround(sc, 1, sc, 0);  // sum of u'^2 over all GPUs + barrier: every slot has landed
    beta = sqrt_(sc[0]);
    if (beta != T(0)) {
      for (int i = rlo + tid; i < rhi; i += nt) wu[i] = wu[i] / beta;
      const T* slots = team.xchg() + (size_t)par * P * npad;
      const T scale_old = first ? T(0) : -beta;
      for (int j = clo + tid; j < chi; j += nt) {
        T acc = T(0);
        for (int q = 0; q < P; ++q) acc += __ldcg(slots + (size_t)q * npad + j);
        wv[j] = wv[j] * scale_old + acc / beta;
      }
    }
    par ^= 1;
    __syncthreads();
  };

  for (int i = clo + tid; i < chi; i += nt) {
    wx[i] = (p.flags & LXB_HAS_Y0) ? p.x[i] : T(0);
    whb[i] = T(0);
    wv[i] = T(0);
  }
  for (int i = rlo + tid; i < rhi; i += nt) wu[i] = b[i];
  __syncthreads();
  const T normb = sqrt_(xsumsq_rows(wu));  // barrier: x visible
  T beta, alpha = T(0);
  {
    // u' = b - A x0 and A^T u' from one read of A
    const T ssq = lsmr_fused_dispatch<T>(ch, A, n, rlo, rhi, wx, wu, wu, T(-1), T(1),
                                         pbuf + (size_t)g.bid * npad, red);
    exchange(ssq, true, beta);
    if (beta != T(0)) alpha = grid_norm2<T>(g, wv, clo, chi, n);
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
    g.sync();  // v complete
    {
      const T ssq = lsmr_fused_dispatch<T>(ch, A, n, rlo, rhi, wv, wu, wu, T(1), -alpha,
                                           pbuf + (size_t)g.bid * npad, red);
      exchange(ssq, false, beta);
    }
    if (beta != T(0)) {
      alpha = grid_norm2<T>(g, wv, clo, chi, n);
      const T den = alpha == T(0) ? T(1) : alpha;
      for (int i = clo + tid; i < chi; i += nt) wv[i] = wv[i] / den;
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
    const T normx = grid_norm2<T>(g, wx, clo, chi, n);
    const T well_posed_tol = p.atol + p.rtol * (normA * normx + normb);
    const T least_squares_tol = p.atol + p.rtol * (normA * normr);
    if (itn >= p.max_steps) istop = 4;
    if (condA > p.conlim) istop = 3;
    if (normAr < least_squares_tol) istop = 2;
    if (normr < well_posed_tol) istop = 1;
  }
  const T normx_final = grid_norm2<T>(g, wx, clo, chi, n);
  int result = krylov_final_result(itn, p.max_steps, p.flags, has_scale);
  if (istop < 3) result = LXB_SUCCESSFUL;
  if (istop == 3) result = LXB_CONLIM;
  for (int i = clo + tid; i < chi; i += nt) p.x[i] = wx[i];
  if (g.bid == 0 && tid == 0) {
    p.result[0] = result;
    p.num_steps[0] = (int32_t)(itn > 2147483647 ? 2147483647 : itn);
    if (p.stats) {
      T* so = p.stats;
      so[0] = T(istop); so[1] = normr; so[2] = normAr; so[3] = sqrt_(normA2);
      so[4] = condA; so[5] = normx_final; so[6] = T(0); so[7] = T(0);
    }
  }
  team.finish(sc);
}

template <typename T>
size_t lsmr_dist_ws_bytes(int m_local, int n) {
  return (grid_part_elems() + pad4(m_local) + (5 + (size_t)grid_blocks()) * pad4(n)) * sizeof(T);
}

template <typename T>
int lsmr_dist_launch(const T* A_local, const T* b_local, T* x, int32_t* result, int32_t* num_steps,
                     T* stats, int m, int m_local, int n, T rtol, T atol, T conlim, int64_t max_steps,
                     int flags, void* ws, size_t ws_bytes, void* const* peers, int world, int rank,
                     cudaStream_t st) {
  if (!A_local || !b_local || !x || !result || !num_steps || !peers || m <= 0 || n <= 0 || m_local < 0 ||
      world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return LXB_E_BADARG;
  constexpr int V = 16 / (int)sizeof(T);
  // the one-read-of-A pass keeps a row's 16-byte chunks in registers: rows must be aligned and short enough
  const int chunks = (n / V + kDistThreads - 1) / kDistThreads;
  if (n % V != 0 || chunks < 1 || chunks > 4) return LXB_E_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(A_local) & 15) return LXB_E_ALIGN;
  if (!ws || ws_bytes < lsmr_dist_ws_bytes<T>(m_local, n)) return LXB_E_WORKSPACE;
  LsmrDistParams<T> dp{};
  dp.k.A = A_local; dp.k.b = b_local; dp.k.x = x; dp.k.result = result; dp.k.num_steps = num_steps;
  dp.k.stats = stats; dp.k.batch = 1; dp.k.m = m_local; dp.k.n = n; dp.k.rtol = rtol; dp.k.atol = atol;
  dp.k.conlim = conlim; dp.k.max_steps = max_steps; dp.k.flags = flags; dp.k.ws = reinterpret_cast<T*>(ws);
  dp.peers = reinterpret_cast<unsigned char* const*>(peers);
  dp.world = world; dp.rank = rank; dp.m_global = m;
  const size_t smem = (96 + kGridMaxK + 8) * sizeof(T);
  auto kern = lsmr_dist_kernel<T>;
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

#define LXB_DEF_LSMR_DIST(sfx, T)                                                                     \
  extern "C" int lxb_lsmr_rowsharded_##sfx(                                                           \
      const T* A_local, const T* b_local, T* x, int32_t* result, int32_t* num_steps, T* stats,       \
      int32_t m, int32_t m_local, int32_t n, T rtol, T atol, T conlim, int64_t max_steps,             \
      int32_t flags, void* workspace, size_t workspace_bytes, void* const* peer_buffers,              \
      int32_t world, int32_t rank, lxb_stream_t stream) {                                             \
    return lxb::lsmr_dist_launch<T>(A_local, b_local, x, result, num_steps, stats, m, m_local, n,     \
                                    rtol, atol, conlim, max_steps, flags, workspace, workspace_bytes, \
                                    peer_buffers, world, rank, (cudaStream_t)stream);                 \
  }                                                                                                   \
  extern "C" size_t lxb_lsmr_rowsharded_workspace_##sfx(int32_t m_local, int32_t n) {                 \
    return lxb::lsmr_dist_ws_bytes<T>(m_local, n);                                                    \
  }                                                                                                   \
  extern "C" size_t lxb_lsmr_rowsharded_symm_bytes_##sfx(int32_t n, int32_t world) {                  \
    return lxb::lsmr_symm_bytes<T>(n, world);                                                         \
  }
LXB_DEF_LSMR_DIST(f32, float)
LXB_DEF_LSMR_DIST(f64, double)

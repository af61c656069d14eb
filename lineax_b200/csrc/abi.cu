// Library-level entry points: version, error strings, launch counter, the
// lineax/_solve.py:104-123 result post-processing kernel and the host-buffer pipeline.
#include "common.cuh"

namespace lxb {

std::atomic<int64_t> g_launch_count{0};

// One warp per system: scan x and b for non-finite entries and rewrite the result code.
template <typename T>
__global__ void __launch_bounds__(256)
    postprocess_kernel(const T* __restrict__ x, int64_t sx, int nx, const T* __restrict__ b,
                       int64_t sb, int nb, int32_t* __restrict__ result, int64_t batch) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t sys = warp; sys < batch; sys += nwarps) {
    bool bad_x = false, bad_b = false;
    for (int i = lane; i < nx; i += 32) bad_x |= !finite_(x[sys * sx + i]);
    for (int i = lane; i < nb; i += 32) bad_b |= !finite_(b[sys * sb + i]);
    bad_x = __any_sync(kFull, bad_x);
    bad_b = __any_sync(kFull, bad_b);
    if (lane == 0) {
      int r = result[sys];
      if (r == LXB_SUCCESSFUL && bad_x) r = LXB_SINGULAR;
      if (r == LXB_SINGULAR && bad_b) r = LXB_NONFINITE_INPUT;
      result[sys] = r;
    }
  }
}

template <typename T>
int postprocess(const T* x, int64_t sx, int nx, const T* b, int64_t sb, int nb, int32_t* result,
                int64_t batch, cudaStream_t st) {
  if (batch < 0 || nx < 0 || nb < 0 || !result || (nx > 0 && !x) || (nb > 0 && !b)) return LXB_E_BADARG;
  if (batch == 0) return 0;
  int64_t blocks = (batch + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  postprocess_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(x, sx, nx, b, sb, nb, result, batch);
  LXB_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace lxb

extern "C" int lxb_version(void) { return 100; }

extern "C" int64_t lxb_launch_count(void) { return lxb::g_launch_count.load(); }

extern "C" const char* lxb_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case LXB_E_BADARG: return "lineax_b200: bad argument (null pointer or negative size)";
    case LXB_E_UNSUPPORTED: return "lineax_b200: shape not supported by the native kernels";
    case LXB_E_WORKSPACE: return "lineax_b200: workspace missing or too small";
    case LXB_E_ALIGN: return "lineax_b200: pointer or stride is not sufficiently aligned";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "lineax_b200: unknown error";
}

#define LXB_DEF_POST(sfx, T)                                                                      \
  extern "C" int lxb_postprocess_##sfx(const T* x, int64_t stride_x, int32_t nx, const T* b,      \
                                       int64_t stride_b, int32_t nb, int32_t* result,             \
                                       int64_t batch, lxb_stream_t stream) {                      \
    return lxb::postprocess<T>(x, stride_x, nx, b, stride_b, nb, result, batch,                   \
                               (cudaStream_t)stream);                                             \
  }
LXB_DEF_POST(f32, float)
LXB_DEF_POST(f64, double)

// ---------------------------------------------------------------- host pipeline ----
extern "C" size_t lxb_host_scratch_bytes(int64_t batch, int32_t n, int32_t elem_bytes) {
  if (batch < 0 || n < 0) return 0;
  return (size_t)batch * ((size_t)n * n + 2 * (size_t)n) * (size_t)elem_bytes + 256;
}

// Chunked, three-stream software pipeline: H2D(A,b) | factor+solve | D2H(x).
extern "C" int lxb_lu_factor_solve_f32_host(const float* A_host, const float* b_host,
                                            float* x_host, int64_t batch, int32_t n,
                                            void* device_scratch, size_t scratch_bytes,
                                            lxb_stream_t stream) {
  if (!A_host || !b_host || !x_host || batch < 0 || n < 0) return LXB_E_BADARG;
  if (batch == 0 || n == 0) return 0;
  if (!device_scratch || scratch_bytes < lxb_host_scratch_bytes(batch, n, 4)) return LXB_E_WORKSPACE;
  cudaStream_t user = (cudaStream_t)stream;
  uintptr_t base = (reinterpret_cast<uintptr_t>(device_scratch) + 255) & ~(uintptr_t)255;
  float* dA = reinterpret_cast<float*>(base);
  float* db = dA + (size_t)batch * n * n;
  float* dx = db + (size_t)batch * n;
  constexpr int kStreams = 3;
  cudaStream_t s[kStreams];
  cudaEvent_t start, done[kStreams];
  LXB_CUDA_TRY(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
  LXB_CUDA_TRY(cudaEventRecord(start, user));
  for (int i = 0; i < kStreams; ++i) {
    LXB_CUDA_TRY(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
    LXB_CUDA_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    LXB_CUDA_TRY(cudaStreamWaitEvent(s[i], start, 0));
  }
  // ~16 MiB of A per chunk keeps copies long enough to saturate the link
  int64_t chunk = (16ll << 20) / ((int64_t)n * n * 4);
  if (chunk < 1) chunk = 1;
  int rc = 0, i = 0;
  for (int64_t off = 0; off < batch && rc == 0; off += chunk, ++i) {
    const int64_t cnt = batch - off < chunk ? batch - off : chunk;
    cudaStream_t st = s[i % kStreams];
    cudaError_t e = cudaMemcpyAsync(dA + off * n * n, A_host + off * n * n,
                                    (size_t)cnt * n * n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(db + off * n, b_host + off * n, (size_t)cnt * n * 4,
                          cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = (int)e; break; }
    rc = lxb_lu_factor_solve_f32(dA + off * n * n, (int64_t)n * n, db + off * n, n, dx + off * n,
                                 nullptr, nullptr, cnt, n, (lxb_stream_t)st);
    if (rc != 0) break;
    e = cudaMemcpyAsync(x_host + off * n, dx + off * n, (size_t)cnt * n * 4,
                        cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = (int)e;
  }
  for (int k = 0; k < kStreams; ++k) {
    cudaEventRecord(done[k], s[k]);
    cudaStreamWaitEvent(user, done[k], 0);
    cudaEventDestroy(done[k]);
    cudaStreamDestroy(s[k]);
  }
  cudaEventDestroy(start);
  return rc;
}

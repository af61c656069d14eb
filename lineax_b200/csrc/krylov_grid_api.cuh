// Host-side policy + declarations for the grid-cooperative Krylov tier (krylov_grid.cu).
#pragma once
#include "krylov_cta.cuh"

namespace lxb {

constexpr int kGridCtasPerSm = 2;
inline int grid_blocks() { return kNumSMs * kGridCtasPerSm; }
constexpr int kGridMaxKHost = 72;
__host__ __device__ inline size_t grid_part_elems() {
  // 2 reduction buffers x kGridMaxK x nb, padded to a multiple of 4 elements
  return ((size_t)2 * kGridMaxKHost * kNumSMs * kGridCtasPerSm + 3) & ~(size_t)3;
}
inline size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }

template <typename T>
size_t cg_grid_ws_bytes(int n) { return (grid_part_elems() + 5 * pad4(n)) * sizeof(T); }
template <typename T>
size_t bicgstab_grid_ws_bytes(int n) { return (grid_part_elems() + 10 * pad4(n)) * sizeof(T); }
template <typename T>
size_t gmres_grid_ws_bytes(int n, int restart) {
  return (grid_part_elems() + (5 + (size_t)restart + 1) * pad4(n)) * sizeof(T);
}
template <typename T>
size_t lsmr_grid_ws_bytes(int m, int n) {
  return (grid_part_elems() + 2 * pad4(m) + (4 + (size_t)grid_blocks()) * pad4(n)) * sizeof(T);
}

// Tier choice: the CTA tier wants >= one system per SM; with fewer systems than half the SMs and
// enough rows to share, the whole grid works on one system at a time.
inline bool use_grid_tier(int64_t batch, int rows, int cols) {
  if (batch <= 0) return false;
  const int64_t work = (int64_t)rows * cols;
  return batch < kNumSMs / 2 && rows >= 256 && work >= (1 << 18);
}

template <typename T>
int cg_grid_launch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st);
template <typename T>
int bicgstab_grid_launch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st);
template <typename T>
int gmres_grid_launch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st);
template <typename T>
int lsmr_grid_launch(KrylovParams<T> p, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace lxb

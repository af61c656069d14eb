/* CPU oracle for batched LU -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Restates the algorithm behind lineax/_solver/lu.py:53,64
 * (jax.scipy.linalg.lu_factor / lu_solve -> LAPACK getrf / getrs on the CPU backend):
 * unblocked right-looking Gaussian elimination with partial pivoting, i.e. reference
 * LAPACK's xGETF2 (pivot = first index of max |a_ik| (ISAMAX), row interchange, column
 * scaled by the reciprocal pivot, rank-1 update), followed by xGETRS
 * (row swaps, unit-lower then upper triangular solve; trans: U^T, L^T, inverse swaps).
 * LAPACK/jaxlib are third-party and absent from /root/reference; the parity anchor for
 * pivots is scipy.linalg.lapack.?getrf (tests/test_oracle_golden.py checks this file
 * against it), and this file fixes the floating-point operation ORDER the CUDA kernels
 * reproduce bit for bit: l = a * (1/pivot); a_ij = fma(-l, u_kj, a_ij) for k ascending;
 * y_i = fma(-l_ik, y_k, y_i); x_k = y_k * (1/u_kk); y_i = fma(-u_ik, x_k, y_i).
 * Build: gcc -O2 -mfma -mavx2 -ffp-contract=off -shared -fPIC (oracle/Makefile); threading is done by the
 * caller (oracle/clib.py splits the batch over a thread pool; ctypes drops the GIL).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_LU(SFX, T, FMA, FABS)                                                             \
  static void getf2_##SFX(T* a, int32_t* piv, int n) {                                           \
    for (int k = 0; k < n; ++k) {                                                                \
      int p = k;                                                                                 \
      T best = FABS(a[(size_t)k * n + k]);                                                       \
      for (int i = k + 1; i < n; ++i) {                                                          \
        T v = FABS(a[(size_t)i * n + k]);                                                        \
        if (v > best || (v != v && best == best)) { best = v; p = i; }                           \
      }                                                                                          \
      piv[k] = p;                                                                                \
      if (p != k)                                                                                \
        for (int j = 0; j < n; ++j) {                                                            \
          T t = a[(size_t)k * n + j];                                                            \
          a[(size_t)k * n + j] = a[(size_t)p * n + j];                                           \
          a[(size_t)p * n + j] = t;                                                              \
        }                                                                                        \
      T r = (T)1 / a[(size_t)k * n + k];                                                         \
      for (int i = k + 1; i < n; ++i) {                                                          \
        T l = a[(size_t)i * n + k] * r;                                                          \
        a[(size_t)i * n + k] = l;                                                                \
        for (int j = k + 1; j < n; ++j)                                                          \
          a[(size_t)i * n + j] = FMA(-l, a[(size_t)k * n + j], a[(size_t)i * n + j]);           \
      }                                                                                          \
    }                                                                                            \
  }                                                                                              \
  static void getrs_##SFX(const T* a, const int32_t* piv, T* y, int n, int trans) {              \
    if (!trans) {                                                                                \
      for (int k = 0; k < n; ++k) { T t = y[k]; y[k] = y[piv[k]]; y[piv[k]] = t; }               \
      for (int k = 0; k < n; ++k)                                                                \
        for (int i = k + 1; i < n; ++i) y[i] = FMA(-a[(size_t)i * n + k], y[k], y[i]);           \
      for (int k = n - 1; k >= 0; --k) {                                                         \
        y[k] = y[k] * ((T)1 / a[(size_t)k * n + k]);                                             \
        for (int i = 0; i < k; ++i) y[i] = FMA(-a[(size_t)i * n + k], y[k], y[i]);               \
      }                                                                                          \
    } else {                                                                                     \
      for (int k = 0; k < n; ++k) {                                                              \
        y[k] = y[k] * ((T)1 / a[(size_t)k * n + k]);                                             \
        for (int i = k + 1; i < n; ++i) y[i] = FMA(-a[(size_t)k * n + i], y[k], y[i]);           \
      }                                                                                          \
      for (int k = n - 1; k >= 0; --k)                                                           \
        for (int i = 0; i < k; ++i) y[i] = FMA(-a[(size_t)k * n + i], y[k], y[i]);               \
      for (int k = n - 1; k >= 0; --k) { T t = y[k]; y[k] = y[piv[k]]; y[piv[k]] = t; }          \
    }                                                                                            \
  }                                                                                              \
  /* lu_factor: A[batch,n,n] -> lu, piv */                                                       \
  void oracle_lu_factor_##SFX(const T* A, T* lu, int32_t* piv, int64_t batch, int n) {           \
    for (int64_t s = 0; s < batch; ++s) {                                                        \
      memcpy(lu + s * n * n, A + s * n * n, sizeof(T) * (size_t)n * n);                          \
      getf2_##SFX(lu + s * n * n, piv + s * n, n);                                               \
    }                                                                                            \
  }                                                                                              \
  void oracle_lu_solve_##SFX(const T* lu, const int32_t* piv, const T* b, T* x, int64_t batch,   \
                             int n, int trans) {                                                 \
    for (int64_t s = 0; s < batch; ++s) {                                                        \
      memcpy(x + s * n, b + s * n, sizeof(T) * (size_t)n);                                       \
      getrs_##SFX(lu + s * n * n, piv + s * n, x + s * n, n, trans);                             \
    }                                                                                            \
  }                                                                                              \
  /* fused factor+solve: the forward substitution is applied during elimination */              \
  void oracle_lu_factor_solve_##SFX(const T* A, const T* b, T* x, T* lu, int32_t* piv,           \
                                    int64_t batch, int n) {                                      \
    oracle_lu_factor_##SFX(A, lu, piv, batch, n);                                                \
    oracle_lu_solve_##SFX(lu, piv, b, x, batch, n, 0);                                           \
  }

DEFINE_LU(f32, float, fmaf, fabsf)
DEFINE_LU(f64, double, fma, fabs)

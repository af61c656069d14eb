// XLA FFI handlers for the lineax_b200 C ABI (include/lineax_b200.h).
//
// This is the "thin XLA FFI layer" of the north star: every handler forwards 1:1 to one C-ABI entry
// point; leading dimensions of the operands are the vmapped batch (`jax.ffi.ffi_call(...,
// vmap_method="broadcast_all")`), scalars travel as FFI attributes, scratch comes from XLA's
// ScratchAllocator, the stream is XLA's.  The Python side that binds these symbols into
// lineax's solver classes is lineax_b200/csrc/xla_ffi/lineax_patch.py.
//
// NOT compiled by default: jaxlib's headers (xla/ffi/api/ffi.h) are absent from this image and JAX is
// not installable here, so this translation unit has never been built.
//   make -C lineax_b200/csrc xla_ffi XLA_FFI_INCLUDE=$(python -c "import jax.ffi; print(jax.ffi.include_dir())")
// builds liblineax_b200_xla.so next to liblineax_b200.so.
#include <cstdint>
#include <string>

#include "xla/ffi/api/ffi.h"

#include "../../../include/lineax_b200.h"

namespace ffi = xla::ffi;

namespace {

template <ffi::DataType DT>
struct Native;
template <>
struct Native<ffi::F32> {
  using type = float;
};
template <>
struct Native<ffi::F64> {
  using type = double;
};

inline ffi::Error Check(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error::Internal(std::string(what) + ": " + lxb_error_string(rc));
}

// [..., r, c] -> number of leading systems
template <typename Buf>
int64_t BatchOf(const Buf& b, int core_dims) {
  auto d = b.dimensions();
  int64_t n = 1;
  for (size_t i = 0; i + core_dims < d.size(); ++i) n *= d[i];
  return n;
}

inline void* Scratch(ffi::ScratchAllocator& alloc, size_t bytes) {
  if (bytes == 0) return nullptr;
  auto p = alloc.Allocate(bytes);
  return p.has_value() ? *p : nullptr;
}

}  // namespace

#define LXB_FFI_FOR_DTYPES(X) X(f32, ffi::F32) X(f64, ffi::F64)

// ---------------------------------------------------------------- LU (lu.py:43-66) ----
#define LXB_FFI_LU(sfx, DT)                                                                                   \
  static ffi::Error LuFactor_##sfx(cudaStream_t stream, ffi::Buffer<DT> a, ffi::ResultBuffer<DT> lu,          \
                                   ffi::ResultBuffer<ffi::S32> piv) {                                         \
    const int32_t n = (int32_t)a.dimensions().back();                                                         \
    return Check(lxb_lu_factor_##sfx(a.typed_data(), (int64_t)n * n, lu->typed_data(), piv->typed_data(),     \
                                     BatchOf(a, 2), n, (lxb_stream_t)stream), "lxb_lu_factor");               \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_lu_factor_##sfx, LuFactor_##sfx,                                      \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Arg<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<ffi::S32>>()); \
  static ffi::Error LuSolve_##sfx(cudaStream_t stream, ffi::Buffer<DT> lu, ffi::Buffer<ffi::S32> piv,         \
                                  ffi::Buffer<DT> b, ffi::ResultBuffer<DT> x, int32_t trans) {                \
    const int32_t n = (int32_t)lu.dimensions().back();                                                        \
    return Check(lxb_lu_solve_##sfx(lu.typed_data(), (int64_t)n * n, piv.typed_data(), n, b.typed_data(), n,  \
                                    x->typed_data(), BatchOf(b, 1), n, trans ? LXB_TRANS : 0,                 \
                                    (lxb_stream_t)stream), "lxb_lu_solve");                                   \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_lu_solve_##sfx, LuSolve_##sfx,                                        \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<DT>>() \
                                    .Ret<ffi::Buffer<DT>>().Attr<int32_t>("trans"));                          \
  static ffi::Error LuFactorSolve_##sfx(cudaStream_t stream, ffi::Buffer<DT> a, ffi::Buffer<DT> b,            \
                                        ffi::ResultBuffer<DT> x) {                                            \
    const int32_t n = (int32_t)a.dimensions().back();                                                         \
    return Check(lxb_lu_factor_solve_##sfx(a.typed_data(), (int64_t)n * n, b.typed_data(), n, x->typed_data(), \
                                           nullptr, nullptr, BatchOf(a, 2), n, (lxb_stream_t)stream),         \
                 "lxb_lu_factor_solve");                                                                      \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_lu_factor_solve_##sfx, LuFactorSolve_##sfx,                           \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>());
LXB_FFI_FOR_DTYPES(LXB_FFI_LU)

// ------------------------------------------------------- Cholesky (cholesky.py:43-78) ----
#define LXB_FFI_CHOL(sfx, DT)                                                                                 \
  static ffi::Error CholFactor_##sfx(cudaStream_t stream, ffi::Buffer<DT> a, ffi::ResultBuffer<DT> f,         \
                                     int32_t nsd) {                                                           \
    const int32_t n = (int32_t)a.dimensions().back();                                                         \
    return Check(lxb_cholesky_factor_##sfx(a.typed_data(), (int64_t)n * n, f->typed_data(), BatchOf(a, 2), n, \
                                           nsd ? LXB_NSD : 0, (lxb_stream_t)stream), "lxb_cholesky_factor");  \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_cholesky_factor_##sfx, CholFactor_##sfx,                              \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Arg<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>().Attr<int32_t>("nsd"));     \
  static ffi::Error CholSolve_##sfx(cudaStream_t stream, ffi::Buffer<DT> f, ffi::Buffer<DT> b,                \
                                    ffi::ResultBuffer<DT> x, int32_t nsd) {                                   \
    const int32_t n = (int32_t)f.dimensions().back();                                                         \
    return Check(lxb_cholesky_solve_##sfx(f.typed_data(), (int64_t)n * n, b.typed_data(), n, x->typed_data(), \
                                          BatchOf(b, 1), n, nsd ? LXB_NSD : 0, (lxb_stream_t)stream),         \
                 "lxb_cholesky_solve");                                                                       \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_cholesky_solve_##sfx, CholSolve_##sfx,                                \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>()     \
                                    .Attr<int32_t>("nsd"));
LXB_FFI_FOR_DTYPES(LXB_FFI_CHOL)

// ------------------------------------------------------------------ QR (qr.py:55-94) ----
#define LXB_FFI_QR(sfx, DT)                                                                                   \
  static ffi::Error QrFactor_##sfx(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<DT> a,     \
                                   ffi::ResultBuffer<DT> aq, ffi::ResultBuffer<DT> taus) {                    \
    auto d = a.dimensions();                                                                                  \
    const int32_t m = (int32_t)d[d.size() - 2], n = (int32_t)d.back();                                        \
    const int64_t batch = BatchOf(a, 2);                                                                      \
    const size_t wsb = lxb_qr_factor_workspace_##sfx(batch, m, n);                                            \
    void* ws = Scratch(scratch, wsb);                                                                         \
    if (wsb && !ws) return ffi::Error::Internal("lxb_qr_factor: scratch allocation failed");                  \
    return Check(lxb_qr_factor_##sfx(a.typed_data(), (int64_t)m * n, aq->typed_data(), taus->typed_data(),    \
                                     batch, m, n, ws, wsb, (lxb_stream_t)stream), "lxb_qr_factor");           \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_qr_factor_##sfx, QrFactor_##sfx,                                      \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Ctx<ffi::ScratchAllocator>().Arg<ffi::Buffer<DT>>()                      \
                                    .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>());                          \
  static ffi::Error QrSolve_##sfx(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<DT> aq,     \
                                  ffi::Buffer<DT> taus, ffi::Buffer<DT> b, ffi::ResultBuffer<DT> x,           \
                                  int32_t trans) {                                                            \
    auto d = aq.dimensions();                                                                                 \
    const int32_t rows = (int32_t)d[d.size() - 2], cols = (int32_t)d.back();                                  \
    const int64_t batch = BatchOf(b, 1);                                                                      \
    const size_t wsb = lxb_qr_solve_workspace_##sfx(batch, rows, cols);                                       \
    void* ws = Scratch(scratch, wsb);                                                                         \
    if (wsb && !ws) return ffi::Error::Internal("lxb_qr_solve: scratch allocation failed");                   \
    return Check(lxb_qr_solve_##sfx(aq.typed_data(), (int64_t)rows * cols, taus.typed_data(), cols,           \
                                    b.typed_data(), trans ? cols : rows, x->typed_data(), batch, rows, cols,  \
                                    trans ? LXB_TRANS : 0, ws, wsb, (lxb_stream_t)stream), "lxb_qr_solve");   \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_qr_solve_##sfx, QrSolve_##sfx,                                        \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Ctx<ffi::ScratchAllocator>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>() \
                                    .Arg<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>().Attr<int32_t>("trans"));
LXB_FFI_FOR_DTYPES(LXB_FFI_QR)

// ------------------------------------------------- Tridiagonal (tridiagonal.py:54-72) ----
#define LXB_FFI_TRIDIAG(sfx, DT)                                                                              \
  static ffi::Error Tridiag_##sfx(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<DT> d,      \
                                  ffi::Buffer<DT> dl, ffi::Buffer<DT> du, ffi::Buffer<DT> b,                  \
                                  ffi::ResultBuffer<DT> x) {                                                  \
    const int32_t n = (int32_t)d.dimensions().back();                                                         \
    const int64_t batch = BatchOf(d, 1);                                                                      \
    const size_t wsb = lxb_tridiagonal_workspace_##sfx(batch, n);                                             \
    void* ws = Scratch(scratch, wsb);                                                                         \
    if (wsb && !ws) return ffi::Error::Internal("lxb_tridiagonal_solve: scratch allocation failed");          \
    return Check(lxb_tridiagonal_solve_##sfx(d.typed_data(), dl.typed_data(), du.typed_data(), n,             \
                                             b.typed_data(), n, x->typed_data(), batch, n, ws, wsb,           \
                                             (lxb_stream_t)stream), "lxb_tridiagonal_solve");                 \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_tridiagonal_solve_##sfx, Tridiag_##sfx,                               \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Ctx<ffi::ScratchAllocator>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>() \
                                    .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>());
LXB_FFI_FOR_DTYPES(LXB_FFI_TRIDIAG)

// -------------------------------------------------------------------- Krylov solvers ----
// outputs: x[..., n], result[...] int32 (RESULTS code), num_steps[...] int32 (+ LSMR stats[..., 8])
#define LXB_FFI_KRYLOV(sfx, DT)                                                                               \
  using T_##sfx = Native<DT>::type;                                                                           \
  static ffi::Error Cg_##sfx(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<DT> a,           \
                             ffi::Buffer<DT> b, ffi::ResultBuffer<DT> x, ffi::ResultBuffer<ffi::S32> result,  \
                             ffi::ResultBuffer<ffi::S32> steps, double rtol, double atol, int32_t max_steps,  \
                             int32_t stabilise_every, int32_t flags) {                                        \
    const int32_t n = (int32_t)a.dimensions().back();                                                         \
    const int64_t batch = BatchOf(b, 1);                                                                      \
    const size_t wsb = lxb_cg_workspace_##sfx(batch, n);                                                      \
    void* ws = Scratch(scratch, wsb);                                                                         \
    if (wsb && !ws) return ffi::Error::Internal("lxb_cg: scratch allocation failed");                         \
    return Check(lxb_cg_##sfx(a.typed_data(), (int64_t)n * n, b.typed_data(), n, nullptr, 0, x->typed_data(), \
                              result->typed_data(), steps->typed_data(), batch, n, (T_##sfx)rtol,             \
                              (T_##sfx)atol, max_steps, stabilise_every, flags, ws, wsb,                      \
                              (lxb_stream_t)stream), "lxb_cg");                                               \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_cg_##sfx, Cg_##sfx,                                                   \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Ctx<ffi::ScratchAllocator>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>() \
                                    .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<ffi::S32>>()                      \
                                    .Ret<ffi::Buffer<ffi::S32>>().Attr<double>("rtol").Attr<double>("atol")   \
                                    .Attr<int32_t>("max_steps").Attr<int32_t>("stabilise_every")              \
                                    .Attr<int32_t>("flags"));                                                 \
  static ffi::Error Bicgstab_##sfx(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<DT> a,     \
                                   ffi::Buffer<DT> b, ffi::ResultBuffer<DT> x,                                \
                                   ffi::ResultBuffer<ffi::S32> result, ffi::ResultBuffer<ffi::S32> steps,     \
                                   double rtol, double atol, int32_t max_steps, int32_t flags) {              \
    const int32_t n = (int32_t)a.dimensions().back();                                                         \
    const int64_t batch = BatchOf(b, 1);                                                                      \
    const size_t wsb = lxb_bicgstab_workspace_##sfx(batch, n);                                                \
    void* ws = Scratch(scratch, wsb);                                                                         \
    if (wsb && !ws) return ffi::Error::Internal("lxb_bicgstab: scratch allocation failed");                   \
    return Check(lxb_bicgstab_##sfx(a.typed_data(), (int64_t)n * n, b.typed_data(), n, nullptr, 0,            \
                                    x->typed_data(), result->typed_data(), steps->typed_data(), batch, n,     \
                                    (T_##sfx)rtol, (T_##sfx)atol, max_steps, flags, ws, wsb,                  \
                                    (lxb_stream_t)stream), "lxb_bicgstab");                                   \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_bicgstab_##sfx, Bicgstab_##sfx,                                       \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Ctx<ffi::ScratchAllocator>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>() \
                                    .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<ffi::S32>>()                      \
                                    .Ret<ffi::Buffer<ffi::S32>>().Attr<double>("rtol").Attr<double>("atol")   \
                                    .Attr<int32_t>("max_steps").Attr<int32_t>("flags"));                      \
  static ffi::Error Gmres_##sfx(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<DT> a,        \
                                ffi::Buffer<DT> b, ffi::ResultBuffer<DT> x, ffi::ResultBuffer<ffi::S32> result, \
                                ffi::ResultBuffer<ffi::S32> steps, double rtol, double atol,                  \
                                int32_t max_steps, int32_t restart, int32_t stagnation_iters, int32_t flags) { \
    const int32_t n = (int32_t)a.dimensions().back();                                                         \
    const int64_t batch = BatchOf(b, 1);                                                                      \
    const size_t wsb = lxb_gmres_workspace_##sfx(batch, n, restart);                                          \
    void* ws = Scratch(scratch, wsb);                                                                         \
    if (wsb && !ws) return ffi::Error::Internal("lxb_gmres: scratch allocation failed");                      \
    return Check(lxb_gmres_##sfx(a.typed_data(), (int64_t)n * n, b.typed_data(), n, nullptr, 0,               \
                                 x->typed_data(), result->typed_data(), steps->typed_data(), batch, n,        \
                                 (T_##sfx)rtol, (T_##sfx)atol, max_steps, restart, stagnation_iters, flags,   \
                                 ws, wsb, (lxb_stream_t)stream), "lxb_gmres");                                \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_gmres_##sfx, Gmres_##sfx,                                             \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Ctx<ffi::ScratchAllocator>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>() \
                                    .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<ffi::S32>>()                      \
                                    .Ret<ffi::Buffer<ffi::S32>>().Attr<double>("rtol").Attr<double>("atol")   \
                                    .Attr<int32_t>("max_steps").Attr<int32_t>("restart")                      \
                                    .Attr<int32_t>("stagnation_iters").Attr<int32_t>("flags"));               \
  static ffi::Error Lsmr_##sfx(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<DT> a,         \
                               ffi::Buffer<DT> b, ffi::ResultBuffer<DT> x, ffi::ResultBuffer<ffi::S32> result, \
                               ffi::ResultBuffer<ffi::S32> steps, ffi::ResultBuffer<DT> stats, double rtol,   \
                               double atol, double conlim, int64_t max_steps, int32_t flags) {                \
    auto d = a.dimensions();                                                                                  \
    const int32_t m = (int32_t)d[d.size() - 2], n = (int32_t)d.back();                                        \
    const int64_t batch = BatchOf(b, 1);                                                                      \
    const size_t wsb = lxb_lsmr_workspace_##sfx(batch, m, n);                                                 \
    void* ws = Scratch(scratch, wsb);                                                                         \
    if (wsb && !ws) return ffi::Error::Internal("lxb_lsmr: scratch allocation failed");                       \
    return Check(lxb_lsmr_##sfx(a.typed_data(), (int64_t)m * n, b.typed_data(), m, x->typed_data(),           \
                                result->typed_data(), steps->typed_data(), stats->typed_data(), batch, m, n,  \
                                (T_##sfx)rtol, (T_##sfx)atol, (T_##sfx)conlim, max_steps, flags, ws, wsb,     \
                                (lxb_stream_t)stream), "lxb_lsmr");                                           \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_lsmr_##sfx, Lsmr_##sfx,                                               \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Ctx<ffi::ScratchAllocator>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>() \
                                    .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<ffi::S32>>()                      \
                                    .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<DT>>().Attr<double>("rtol") \
                                    .Attr<double>("atol").Attr<double>("conlim").Attr<int64_t>("max_steps")   \
                                    .Attr<int32_t>("flags"));
LXB_FFI_FOR_DTYPES(LXB_FFI_KRYLOV)

// ------------------------------------- result rewriting, _solve.py:104-123 (in place on `result`) ----
#define LXB_FFI_POST(sfx, DT)                                                                                 \
  static ffi::Error Post_##sfx(cudaStream_t stream, ffi::Buffer<DT> x, ffi::Buffer<DT> b,                     \
                               ffi::Buffer<ffi::S32> result_in, ffi::ResultBuffer<ffi::S32> result) {         \
    const int32_t nx = (int32_t)x.dimensions().back(), nb = (int32_t)b.dimensions().back();                   \
    const int64_t batch = BatchOf(x, 1);                                                                      \
    if (result->typed_data() != result_in.typed_data())                                                       \
      cudaMemcpyAsync(result->typed_data(), result_in.typed_data(), sizeof(int32_t) * batch,                  \
                      cudaMemcpyDeviceToDevice, stream);                                                      \
    return Check(lxb_postprocess_##sfx(x.typed_data(), nx, nx, b.typed_data(), nb, nb, result->typed_data(),  \
                                       batch, (lxb_stream_t)stream), "lxb_postprocess");                      \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(lxb_ffi_postprocess_##sfx, Post_##sfx,                                        \
                                ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<ffi::S32>>() \
                                    .Ret<ffi::Buffer<ffi::S32>>());
LXB_FFI_FOR_DTYPES(LXB_FFI_POST)

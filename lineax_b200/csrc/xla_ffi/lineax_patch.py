"""Bind the lineax_b200 XLA-FFI handlers into patrick-kidger/lineax (JAX host).

NOT importable in this image (jax / jaxlib / equinox / lineax are absent) and
therefore never executed here; it is the reference-side binding a lineax maintainer would add, kept next
to the handler sources (lxb_xla_ffi.cc) it binds.  `install()` registers every handler as an FFI target
and replaces the bodies of `init` / `compute` of the eight solver classes for MATERIALISED operators on a
CUDA device; lineax's custom_jvp / transpose / vmap rules (`lineax/_solve.py:151-332`) keep composing
because they only re-bind the primitive with transposed / conjugated state.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "..", "..", "liblineax_b200_xla.so")

_TARGETS = ("lu_factor", "lu_solve", "lu_factor_solve", "cholesky_factor", "cholesky_solve", "qr_factor",
            "qr_solve", "tridiagonal_solve", "cg", "bicgstab", "gmres", "lsmr", "postprocess")


def _sfx(dtype):
    import jax.numpy as jnp

    return {jnp.dtype("float32"): "f32", jnp.dtype("float64"): "f64"}[jnp.dtype(dtype)]


def install():
    import jax
    import jax.numpy as jnp
    import lineax as lx
    from lineax._solution import RESULTS
    from lineax._solver.misc import ravel_vector, unravel_solution

    lib = ctypes.CDLL(_LIB)
    for name in _TARGETS:
        for sfx in ("f32", "f64"):
            jax.ffi.register_ffi_target(f"lxb_{name}_{sfx}", jax.ffi.pycapsule(getattr(lib, f"lxb_ffi_{name}_{sfx}")),
                                        platform="CUDA")

    def call(name, outs, *args, **attrs):
        # leading batch dimensions are consumed natively by the kernels: declared batching rule
        return jax.ffi.ffi_call(name, outs, vmap_method="broadcast_all")(*args, **attrs)

    # ---- LU (lineax/_solver/lu.py:43-66)
    def lu_init(self, operator, options):
        del options
        matrix = operator.as_matrix()
        n = matrix.shape[-1]
        sfx = _sfx(matrix.dtype)
        lu, piv = call(f"lxb_lu_factor_{sfx}",
                       (jax.ShapeDtypeStruct(matrix.shape, matrix.dtype), jax.ShapeDtypeStruct((n,), jnp.int32)), matrix)
        from lineax._solver.misc import pack_structures
        import equinox.internal as eqxi

        return (lu, piv), pack_structures(operator), eqxi.Static(False)

    def lu_compute(self, state, vector, options):
        del options
        (lu, piv), packed_structures, transpose = state
        vector = ravel_vector(vector, packed_structures)
        x = call(f"lxb_lu_solve_{_sfx(lu.dtype)}", jax.ShapeDtypeStruct(vector.shape, lu.dtype), lu, piv, vector,
                 trans=np.int32(transpose.value))
        return unravel_solution(x, packed_structures), RESULTS.successful, {}

    lx.LU.init, lx.LU.compute = lu_init, lu_compute

    # ---- CG (lineax/_solver/cg.py:114-227); BiCGStab / GMRES / LSMR follow the same pattern
    def cg_compute(self, state, vector, options):
        operator, is_nsd = state
        if options.get("preconditioner") is not None or options.get("y0") is not None:
            return _orig_cg_compute(self, state, vector, options)  # options stay on the JAX path
        matrix = operator.as_matrix()
        n = matrix.shape[-1]
        flat, unravel = jax.flatten_util.ravel_pytree(vector)
        max_steps = 10 * n if self.max_steps is None else self.max_steps
        flags = (2 if is_nsd.value else 0) | (0 if self.max_steps is None else 4)
        x, result, steps = call(
            f"lxb_cg_{_sfx(matrix.dtype)}",
            (jax.ShapeDtypeStruct(flat.shape, matrix.dtype), jax.ShapeDtypeStruct((), jnp.int32),
             jax.ShapeDtypeStruct((), jnp.int32)),
            matrix, flat, rtol=float(self.rtol), atol=float(self.atol), max_steps=np.int32(max_steps),
            stabilise_every=np.int32(self.stabilise_every or 0), flags=np.int32(flags))
        return unravel(x), RESULTS.promote(result), {"num_steps": steps, "max_steps": self.max_steps}

    _orig_cg_compute = lx.CG.compute
    lx.CG.compute = cg_compute
    return lib

"""JAX front end for libipx.so: interp3d's cubic branch as an XLA custom call with a
custom_jvp rule, the way it would be patched into interpax.

NOT EXERCISED IN THIS REPOSITORY'S ENVIRONMENT: jax / jaxlib are not installed in the build or
GPU image and cannot be installed (no wheels, no network), and the XLA FFI shim
(csrc/xla_ffi_shim.cc) needs jaxlib's headers to compile.  Importing this module without JAX
raises ImportError with that explanation; nothing else in the package imports it.  The tested
path to the same kernels is interpax_b200.api (ctypes + torch).
"""

from __future__ import annotations

import ctypes
import os

try:  # pragma: no cover - JAX is absent here
    import jax
    import jax.numpy as jnp
    import numpy as np
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "interpax_b200.jax_frontend needs jax/jaxlib (absent in this environment); use "
        "interpax_b200.api, which drives the same C ABI through ctypes"
    ) from e

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "libipx_xla.so")  # built from csrc/xla_ffi_shim.cc where JAX exists


def _register():  # pragma: no cover
    lib = ctypes.CDLL(_SHIM)
    jax.ffi.register_ffi_target("ipx_eval3d_f64", jax.ffi.pycapsule(lib.IpxEval3dF64), platform="CUDA")


def _isbool(x):
    return isinstance(x, bool) or (hasattr(x, "dtype") and x.dtype == bool)


def params_bytes(derivative, extrap6, periods):  # pragma: no cover
    """ipx_params (include/ipx.h) as a traced uint8 vector: derivative, extrap and period stay
    traced values, exactly as in interpax where only `method` is static under jit."""
    rows = []
    for ax in range(3):
        lo, hi = extrap6[2 * ax], extrap6[2 * ax + 1]
        keep = lambda e: jnp.asarray(e).astype(jnp.int32) if _isbool(e) else jnp.int32(0)
        fill = lambda e: jnp.float64(jnp.nan) if _isbool(e) else jnp.asarray(e, jnp.float64)
        ints = jnp.stack([jnp.asarray(derivative[ax], jnp.int32), keep(lo), keep(hi), jnp.int32(0)])
        per = 0.0 if periods[ax] is None else periods[ax]
        dbls = jnp.stack([fill(lo), fill(hi), jnp.abs(jnp.asarray(per, jnp.float64))])
        rows += [jax.lax.bitcast_convert_type(ints, jnp.uint8).ravel(),
                 jax.lax.bitcast_convert_type(dbls, jnp.uint8).ravel()]
    return jnp.concatenate(rows)


def eval3d(xq, yq, zq, x, y, z, tabs, params, perms, method_class=2, periodic_mask=0):  # pragma: no cover
    """One custom call = the body of interp3d's cubic branch (interpax/_spline.py:1051-1103)."""
    batch = int(np.prod(tabs[0].shape[3:]))
    out = jax.ShapeDtypeStruct((xq.shape[0], batch), jnp.float64)
    return jax.ffi.ffi_call("ipx_eval3d_f64", out, vmap_method="broadcast_all")(
        xq, yq, zq, x, y, z, *tabs, params, *perms,
        method_class=np.int32(method_class), periodic_mask=np.int32(periodic_mask), batch=np.int64(batch))

"""JAX front end for libipx.so: interp1d / interp2d / interp3d, Interpolator1D/2D/3D and
approx_df with interpax's signatures, every numerical body an XLA custom call (jax.ffi) into
the C ABI of include/ipx.h, with the differentiation and batching rules the reference gets from
tracing jax.numpy:

  jit    derivative orders, extrap flags / fill values and periods stay TRACED operands (only
         `method` is static, interpax/_spline.py:444,595,808): they travel as an ipx_params
         struct in a device buffer, built with jnp ops (`params_buffer`);
  grad / jvp with respect to the QUERIES: custom_jvp -- the tangent is the same call with that
         axis' derivative order + 1 times the query tangent (zero where an extrapolation fill
         replaced the value), exactly d/dxq of the piecewise polynomial;
  grad / jvp with respect to the TABLES (f, fx, ...): the evaluation is linear in them, so the
         tangent is the same call on the tangent tables and the transpose is ipx_vjp_tables_*
         (`jax.custom_derivatives.linear_call`);
  vmap   `vmap_method="expand_dims"`: batched queries extend the query count, batched tables are
         looped over in the handler (csrc/xla_ffi_shim.cc).

NOT EXERCISED IN THIS REPOSITORY'S ENVIRONMENT: jax / jaxlib are not installed in the build or
GPU image and cannot be installed (no wheels, no network), and the XLA FFI shim needs jaxlib's
headers to compile.  Importing this module without JAX raises ImportError with that
explanation; nothing else in the package imports it.  The tested path to the same kernels is
interpax_b200.api (ctypes + torch), whose argument handling this file mirrors.
"""

from __future__ import annotations

import ctypes
import os
from functools import partial

try:  # pragma: no cover - JAX is absent here
    import jax
    import jax.numpy as jnp
    import numpy as np
    from jax.custom_derivatives import linear_call
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "interpax_b200.jax_frontend needs jax/jaxlib (absent in this environment); use "
        "interpax_b200.api, which drives the same C ABI through ctypes"
    ) from e

# everything below is unexercised here (see the module docstring)
# pragma: no cover

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "libipx_xla.so")  # built from csrc/xla_ffi_shim.cc where JAX exists
_CORE = os.path.join(_HERE, "libipx.so")

IPX_NEAREST, IPX_LINEAR, IPX_CUBIC = 0, 1, 2
IPX_SOA, IPX_AOS = 0, 1
SLOPE = {"cubic": 0, "cubic2": 1, "cardinal": 2, "catmull-rom": 2, "monotonic": 3, "monotonic-0": 4, "akima": 5,
         "nearest": 6, "linear": 6}
STENCIL = {"cubic": 0, "cardinal": 2, "catmull-rom": 2}  # methods evaluated straight from f
CUBIC_METHODS = ("cubic", "cubic2", "cardinal", "catmull-rom", "akima", "monotonic", "monotonic-0")
TABLES = {1: ("f", "fx"), 2: ("f", "fx", "fy", "fxy"), 3: ("f", "fx", "fy", "fz", "fxy", "fxz", "fyz", "fxyz")}
CHAIN = {1: (("fx", "f", 0),), 2: (("fx", "f", 0), ("fy", "f", 1), ("fxy", "fx", 1)),
         3: (("fx", "f", 0), ("fy", "f", 1), ("fz", "f", 2), ("fxy", "fx", 1), ("fxz", "fx", 2), ("fyz", "fy", 2),
             ("fxyz", "fxy", 2))}
_HANDLERS = ("IpxEval", "IpxEvalStencil", "IpxHaloTable", "IpxSlopes", "IpxSlopesPack", "IpxPack",
             "IpxPeriodicKnots", "IpxPadPeriodic", "IpxVjpTables", "IpxBinIndex", "IpxSlopesVjp")
_registered = False


def register():
    """Load the shim and register its handlers as XLA custom-call targets (platform CUDA)."""
    global _registered
    if _registered:
        return
    ctypes.CDLL(_CORE, mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(_SHIM)
    for name in _HANDLERS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    _registered = True


def _core():
    lib = ctypes.CDLL(_CORE)
    lib.ipx_halo_table_elems.restype = ctypes.c_int64
    lib.ipx_slopes_workspace_bytes.restype = ctypes.c_size_t
    lib.ipx_slopes_vjp_workspace_bytes.restype = ctypes.c_size_t
    lib.ipx_slopes_vjp_workspace_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                                   ctypes.c_int32]
    return lib


# ------------------------------------------------------------------------------ argument parsing
def _isbool(x):
    return isinstance(x, bool) or (hasattr(x, "dtype") and x.dtype == bool)


def _parse_ndarg(arg, n):
    """interpax/_spline.py:1154-1161."""
    try:
        k = len(arg)
    except TypeError:
        return tuple(arg for _ in range(n))
    assert k == n, "got too many args"
    return tuple(arg)


def _parse_extrap(extrap, n):
    """interpax/_spline.py:1164-1177."""
    if _isbool(extrap) or jnp.ndim(extrap) == 0:
        return tuple(extrap for _ in range(2 * n))
    if len(extrap) == 2 and jnp.ndim(extrap[0]) == 0:
        return tuple(e for _ in range(n) for e in extrap)
    if len(extrap) == n and all(len(extrap[i]) == 2 for i in range(n)):
        return tuple(eij for ei in extrap for eij in ei)
    raise ValueError("extrap should either be a scalar, 2 element sequence (lo, hi), "
                     "or a sequence with 2 elements for each dimension")


def params_buffer(ndim, derivative, extrap6, periods, zero_fill=False):
    """ipx_params (include/ipx.h: per axis 4 x int32 then 3 x float64, three axes = 120 bytes) as a
    traced uint8 vector, so derivative / extrap / period remain run-time values under jit.
    ``zero_fill``: every fill value becomes 0 (the tangent of a constant fill, NaN included)."""
    rows = []
    for ax in range(3):
        if ax < ndim:
            lo, hi = extrap6[2 * ax], extrap6[2 * ax + 1]
            keep = lambda e: jnp.asarray(e).astype(jnp.int32) if _isbool(e) else jnp.int32(0)
            fill = lambda e: (jnp.float64(0.0) if zero_fill else
                              jnp.float64(jnp.nan) if _isbool(e) else jnp.asarray(e, jnp.float64))
            ints = jnp.stack([jnp.asarray(derivative[ax], jnp.int32), keep(lo), keep(hi), jnp.int32(0)])
            per = 0.0 if periods[ax] is None else periods[ax]
            dbls = jnp.stack([fill(lo), fill(hi), jnp.abs(jnp.asarray(per, jnp.float64))])
        else:
            ints, dbls = jnp.zeros(4, jnp.int32), jnp.zeros(3, jnp.float64)
        rows += [jax.lax.bitcast_convert_type(ints, jnp.uint8).ravel(),
                 jax.lax.bitcast_convert_type(dbls, jnp.uint8).ravel()]
    return jnp.concatenate(rows)


def _empty(dtype):
    return jnp.zeros((0,), dtype)


# ------------------------------------------------------------------------------ custom calls
def _call(name, out_types, *operands, **attrs):
    register()
    return jax.ffi.ffi_call(name, out_types, vmap_method="expand_dims")(*operands, **attrs)


def approx_df(x, f, method="cubic", axis=-1, **kwargs):
    """interpax.approx_df (interpax/_fd_derivs.py:9-76) as one custom call (ipx_slopes_*)."""
    if method not in SLOPE:
        raise ValueError(f"got unknown method {method}")
    f = jnp.asarray(f)
    x = jnp.asarray(x)
    dt = jnp.result_type(f, x)
    if jnp.issubdtype(dt, jnp.complexfloating):
        if method == "akima":
            raise NotImplementedError("complex akima: use interpax_b200.api (pair-aware kernel)")
        return approx_df(x, f.real, method, axis, **kwargs) + 1j * approx_df(x, f.imag, method, axis, **kwargs)
    f = f.astype(dt)
    x = x.astype(dt)
    axis = axis % f.ndim
    n = f.shape[axis]
    bc = kwargs.get("bc_type", "not-a-knot")
    bc = (bc, bc) if isinstance(bc, str) else tuple(bc)
    kinds, vals = [], []
    for b in bc:
        if isinstance(b, str):
            if b == "periodic":
                raise NotImplementedError
            kinds.append({"not-a-knot": 0, "clamped": 1, "natural": 2}[b])
            vals.append(_empty(dt))
        else:
            kinds.append(1 if b[0] == 1 else 2)
            vals.append(jnp.asarray(b[1], dt).ravel())
    ws = int(_core().ipx_slopes_workspace_bytes(SLOPE[method], n, jnp.dtype(dt).itemsize))
    attrs = dict(method=np.int32(SLOPE[method]), axis=np.int32(axis),
                 c=np.float64(kwargs.get("c", 0.0) if method == "cardinal" else 0.0),
                 bc_start=np.int32(kinds[0]), bc_end=np.int32(kinds[1]))

    def forward(res, ff):
        xx, v0, v1 = res
        return _call("IpxSlopes",
                     (jax.ShapeDtypeStruct(f.shape, dt), jax.ShapeDtypeStruct((max(ws, 16),), jnp.uint8)),
                     xx, ff, v0, v1, **attrs)[0]

    if method not in ("cubic", "cubic2", "cardinal", "catmull-rom"):
        return forward((x, vals[0], vals[1]), f)  # monotonic / akima: not linear in f, no AD rule here

    # linear in f (plus a constant from the boundary values): reverse mode = ipx_slopes_vjp_*
    # (transposed stencils; cubic2: transposed tridiagonal solve).  The boundary-value constant is
    # added outside the linear call.
    def transpose(res, ct):
        xx = res[0]
        wsv = int(_core().ipx_slopes_vjp_workspace_bytes(SLOPE[method], n, int(np.prod(f.shape[:axis], dtype=np.int64)),
                                                         int(np.prod(f.shape[axis + 1:], dtype=np.int64)),
                                                         jnp.dtype(dt).itemsize))
        return _call("IpxSlopesVjp",
                     (jax.ShapeDtypeStruct(f.shape, dt), jax.ShapeDtypeStruct((max(wsv, 16),), jnp.uint8)),
                     xx, ct, **attrs)[0]

    zeros = (x, jnp.zeros_like(vals[0]), jnp.zeros_like(vals[1]))
    lin = linear_call(forward, transpose, zeros, f)
    if vals[0].size or vals[1].size:
        const = jax.lax.stop_gradient(forward((x, vals[0], vals[1]), jnp.zeros_like(f)))
        return lin + const
    return lin


def periodic_knots(x, period):
    """Knot side of _make_periodic (interpax/_spline.py:1117-1122): (xpad[n+2], perm[n])."""
    n = x.shape[0]
    xpad, perm, _ = _call("IpxPeriodicKnots",
                          (jax.ShapeDtypeStruct((n + 2,), x.dtype), jax.ShapeDtypeStruct((n,), jnp.int32),
                           jax.ShapeDtypeStruct(((n + 4) * x.dtype.itemsize,), jnp.uint8)),
                          x, period=np.float64(abs(float(period))))
    return xpad, perm


def pad_periodic(t, axis, perm):
    shape = list(t.shape)
    shape[axis] += 2
    return _call("IpxPadPeriodic", jax.ShapeDtypeStruct(tuple(shape), t.dtype), t, perm, axis=np.int32(axis))


def pack_records(ndim, tabs):
    f = tabs[0]
    return _call("IpxPack", jax.ShapeDtypeStruct(f.shape + (1 << ndim,), f.dtype), *tabs, ndim=np.int32(ndim))


def halo_table(ndim, f, mask, perms):
    narr = (ctypes.c_int32 * ndim)(*[int(v) for v in f.shape[:ndim]])
    elems = int(_core().ipx_halo_table_elems(ndim, narr, int(mask), jnp.dtype(f.dtype).itemsize))
    return _call("IpxHaloTable", jax.ShapeDtypeStruct((elems,), f.dtype), f,
                 *[p if p is not None else _empty(jnp.int32) for p in perms], ndim=np.int32(ndim),
                 periodic_mask=np.int32(mask))


# ------------------------------------------------------------------------------ evaluation + AD rules
def _eval_tables(static, res, tables):
    """out = E(queries, knots, params) . tables -- LINEAR in `tables` (tuple of 1 or 2^d arrays)."""
    ndim, mclass, mask, layout = static
    qs, knots, perms, params = res
    batch = int(np.prod(tables[0].shape[ndim:])) if layout == IPX_SOA else int(np.prod(tables[0].shape[ndim:-1]))
    out = jax.ShapeDtypeStruct(qs[0].shape + (batch,), tables[0].dtype)
    return _call("IpxEval", out, *qs, *knots, *perms, params, *tables, ndim=np.int32(ndim), layout=np.int32(layout),
                 method_class=np.int32(mclass), periodic_mask=np.int32(mask), batch=np.int64(batch),
                 ntab=np.int32(len(tables)))


def _eval_tables_transpose(static, res, ct):
    """Transpose of _eval_tables with respect to the tables: ipx_vjp_tables_* (SoA tables)."""
    ndim, mclass, mask, layout = static
    qs, knots, perms, params, shapes = res
    ntab = len(shapes)
    batch = int(np.prod(shapes[0][ndim:]))
    outs = tuple(jax.ShapeDtypeStruct(s, ct.dtype) for s in shapes)
    n = [int(v) for v in shapes[0][:ndim]]
    return tuple(_call("IpxVjpTables", outs, *qs, *knots, *perms, params, ct, ndim=np.int32(ndim),
                       method_class=np.int32(mclass), periodic_mask=np.int32(mask), batch=np.int64(batch),
                       ntab=np.int32(ntab), n=np.asarray(n, np.int64)))


def _linear_eval(static, qs, knots, perms, params, tables):
    """The evaluation as a linear map of the tables with its transpose attached, so that reverse
    mode with respect to f / fx / ... works through the opaque custom call."""
    if static[3] != IPX_SOA:  # packed records are an internal layout: not differentiated through
        return _eval_tables(static, (qs, knots, perms, params), tables)
    shapes = tuple(t.shape for t in tables)
    return linear_call(lambda r, t: _eval_tables(static, r[:4], t),
                       lambda r, ct: _eval_tables_transpose(static, r, ct),
                       (qs, knots, perms, params, shapes), tables)


@partial(jax.custom_jvp, nondiff_argnums=(0,))
def _eval(static, qs, knots, perms, derivs, extrap6, periods, tables):
    ndim, zero_fill = static[0], static[4]
    return _linear_eval(static[:4], qs, knots, perms, params_buffer(ndim, derivs, extrap6, periods, zero_fill),
                        tables)


@_eval.defjvp
def _eval_jvp(static, primals, tangents):
    qs, knots, perms, derivs, extrap6, periods, tables = primals
    dqs, _, _, _, _, _, dtables = tangents
    ndim = static[0]
    zero = jax.custom_derivatives.SymbolicZero
    tstatic = static[:4] + (True,)  # tangents of filled (constant) results are zero
    out = _eval(static, qs, knots, perms, derivs, extrap6, periods, tables)
    tan = jnp.zeros_like(out)
    for ax in range(ndim):
        if type(dqs[ax]) is zero:
            continue
        # d out / d q_ax: one more derivative along that axis (the piecewise polynomial's own)
        d1 = tuple(d + (1 if k == ax else 0) for k, d in enumerate(derivs))
        g = _eval(tstatic, qs, knots, perms, d1, extrap6, periods, tables)
        tan = tan + g * dqs[ax].reshape(dqs[ax].shape + (1,))
    if not all(type(t) is zero for t in dtables):
        dt = tuple(jnp.zeros_like(p) if type(t) is zero else t for p, t in zip(tables, dtables))
        tan = tan + _eval(tstatic, qs, knots, perms, derivs, extrap6, periods, dt)
    return out, tan


def _eval_stencil(ndim, qs, knots, sknots, ns, halo, records, method, c, mask, params):
    out, _ = _call("IpxEvalStencil",
                   (jax.ShapeDtypeStruct(qs[0].shape + (1,), halo.dtype), jax.ShapeDtypeStruct((4,), jnp.int32)),
                   *qs, *knots, *[s if s is not None else _empty(halo.dtype) for s in sknots], params, halo,
                   records if records is not None else _empty(halo.dtype), ndim=np.int32(ndim),
                   slope_method=np.int32(STENCIL[method]), c=np.float64(c), periodic_mask=np.int32(mask),
                   hint=np.int32(0), n=np.asarray(ns, np.int64))
    return out


# ------------------------------------------------------------------------------ public functions
def _interp(ndim, qs, knots, f, method, derivative, extrap, period, kwargs):
    """Body shared by interp1d/2d/3d (interpax/_spline.py:503-592, 663-805, 882-1105)."""
    names = TABLES[ndim]
    given = {n: kwargs.pop(n, None) for n in names[1:]}
    kwargs.pop("axis", None)
    xt = jnp.result_type(*knots, *qs)
    f = jnp.asarray(f)
    ft = jnp.result_type(f, xt)
    if jnp.issubdtype(ft, jnp.complexfloating):  # the evaluation is real-linear in f
        kw = dict(kwargs)
        re = _interp(ndim, qs, knots, f.real, method, derivative, extrap, period, dict(kw))
        im = _interp(ndim, qs, knots, f.imag, method, derivative, extrap, period, dict(kw))
        return re + 1j * im
    qs = jnp.broadcast_arrays(*[jnp.asarray(q, ft) for q in qs])
    qshape = qs[0].shape
    qs = tuple(q.reshape(-1) for q in qs)
    knots = tuple(jnp.asarray(k, ft) for k in knots)
    f = f.astype(ft)
    for ax, name in zip(range(ndim), "xyz"):
        if knots[ax].ndim != 1 or knots[ax].shape[0] != f.shape[ax]:
            raise ValueError(f"{name} and f must be arrays of equal length")
    periods = _parse_ndarg(period, ndim)
    derivs = tuple(_parse_ndarg(derivative, ndim))
    ex = _parse_extrap(extrap, ndim)
    mclass = IPX_NEAREST if method == "nearest" else IPX_LINEAR if method == "linear" else IPX_CUBIC
    tabs = {"f": f}
    tabs.update({n: None if v is None else jnp.asarray(v, ft) for n, v in given.items()})
    mask = 0
    perms = [_empty(jnp.int32)] * ndim
    kuse = list(knots)
    pk = [periodic_knots(knots[ax], periods[ax]) if periods[ax] is not None else None for ax in range(ndim)]
    missing = [n for n in names[1:] if tabs[n] is None]
    if mclass == IPX_CUBIC and missing and any(p is not None for p in pk):
        # the functional form differentiates the PADDED tables (_spline.py:924-938 then :1027-1040)
        for ax in range(ndim):
            if pk[ax] is not None:
                for n in names:
                    if tabs[n] is not None:
                        tabs[n] = pad_periodic(tabs[n], ax, pk[ax][1])
                kuse[ax] = pk[ax][0]
                mask |= 1 << (4 + ax)
    else:
        for ax in range(ndim):
            if pk[ax] is not None:
                kuse[ax], perms[ax] = pk[ax]
                mask |= 1 << ax
    if mclass == IPX_CUBIC:
        for name, src, ax in CHAIN[ndim]:
            if tabs[name] is None:
                tabs[name] = approx_df(kuse[ax], tabs[src], method, ax, **kwargs)
        tables = tuple(tabs[n] for n in names)
    else:
        tables = (tabs["f"],)
    out = _eval((ndim, mclass, mask, IPX_SOA, False), qs, tuple(kuse), tuple(perms), derivs, ex, tuple(periods),
                tables)
    return out.reshape(qshape + f.shape[ndim:])


def interp1d(xq, x, f, method="cubic", derivative=0, extrap=False, period=None, **kwargs):
    """interpax.interp1d (interpax/_spline.py:445-592)."""
    return _interp(1, (xq,), (x,), f, method, derivative, extrap, period, dict(kwargs))


def interp2d(xq, yq, x, y, f, method="cubic", derivative=0, extrap=False, period=None, **kwargs):
    """interpax.interp2d (interpax/_spline.py:596-805)."""
    return _interp(2, (xq, yq), (x, y), f, method, derivative, extrap, period, dict(kwargs))


def interp3d(xq, yq, zq, x, y, z, f, method="cubic", derivative=0, extrap=False, period=None, **kwargs):
    """interpax.interp3d (interpax/_spline.py:809-1105)."""
    return _interp(3, (xq, yq, zq), (x, y, z), f, method, derivative, extrap, period, dict(kwargs))


# ------------------------------------------------------------------------------ containers
@jax.tree_util.register_pytree_node_class
class _InterpolatorND:
    """Interpolator1D/2D/3D (interpax/_spline.py:48-441) as a pytree.  Leaves: knots, f, the slope
    tables (`derivs`, as in the reference) and -- so that __call__ reaches the fast kernels -- the
    halo copy of f (cubic / cardinal / catmull-rom on scalar tables: evaluation straight from f)
    and the packed records (tile kernel).  `method`, `extrap`'s structure and `period`'s structure
    are static."""

    _ndim = 0

    def __init__(self, *args, method="cubic", extrap=False, period=None, **kwargs):
        ndim = self._ndim
        knots, f = args[:ndim], jnp.asarray(args[ndim])
        ft = jnp.result_type(f, *knots)
        self.knots = tuple(jnp.asarray(k, ft) for k in knots)
        self.f = f.astype(ft)
        self.method, self.extrap, self.period = method, extrap, period
        names = TABLES[ndim]
        derivs = {n: kwargs.pop(n, None) for n in names[1:]}
        self.kwargs = kwargs
        tabs = {"f": self.f}
        for name, src, ax in CHAIN[ndim]:
            tabs[name] = derivs[name] if derivs[name] is not None else approx_df(self.knots[ax], tabs[src], method,
                                                                                 ax, **kwargs)
        self.derivs = {n: tabs[n] for n in names[1:]}
        self.records = self.halo = None
        scalar = self.f.ndim == ndim and not jnp.issubdtype(ft, jnp.complexfloating)
        if method in CUBIC_METHODS and scalar:
            self.records = pack_records(ndim, [tabs[n] for n in names])
            if method in STENCIL and all(v is None for v in derivs.values()):
                pmask = sum(1 << ax for ax, p in enumerate(_parse_ndarg(period, ndim)) if p is not None)
                self.halo = halo_table(ndim, self.f, pmask, [None] * ndim)

    def tree_flatten(self):
        leaves = (self.knots, self.f, self.derivs, self.records, self.halo, self.extrap, self.period)
        return leaves, (self.method, tuple(self.kwargs.items()))

    @classmethod
    def tree_unflatten(cls, aux, leaves):
        obj = object.__new__(cls)
        obj.knots, obj.f, obj.derivs, obj.records, obj.halo, obj.extrap, obj.period = leaves
        obj.method, kw = aux
        obj.kwargs = dict(kw)
        return obj

    def _call(self, qs, derivative):
        ndim = self._ndim
        if self.halo is None:
            fun = (interp1d, interp2d, interp3d)[ndim - 1]
            return fun(*qs, *self.knots, self.f, self.method, derivative, self.extrap, self.period, **self.derivs,
                       **self.kwargs)
        return _stencil_call(self, tuple(jnp.asarray(q, self.f.dtype) for q in qs), tuple(derivative), False)

    def __call__(self, *args):
        ndim = self._ndim
        qs = args[:ndim]
        d = tuple(args[ndim:]) + (0,) * (2 * ndim - len(args))
        return self._call(qs, d)


@partial(jax.custom_jvp, nondiff_argnums=(2, 3))
def _stencil_call(ip, qs, derivative, zero_fill):
    """Interpolator*.__call__ on the halo table (+ records): the kernels pick the faster form on
    the device.  Differentiable with respect to the queries (derivative + 1); with respect to the
    tables use the functional form, whose evaluation carries the transpose rule."""
    ndim = ip._ndim
    qs = jnp.broadcast_arrays(*qs)
    shape = qs[0].shape
    periods = _parse_ndarg(ip.period, ndim)
    ex = _parse_extrap(ip.extrap, ndim)
    mask, kuse = 0, list(ip.knots)
    for ax in range(ndim):
        if periods[ax] is not None:
            kuse[ax], _ = periodic_knots(ip.knots[ax], periods[ax])
            mask |= 1 << ax
    c = float(ip.kwargs.get("c", 0.0)) if ip.method == "cardinal" else 0.0
    out = _eval_stencil(ndim, tuple(q.reshape(-1) for q in qs), tuple(kuse), ip.knots, ip.f.shape[:ndim], ip.halo,
                        ip.records, ip.method, c, mask, params_buffer(ndim, derivative, ex, periods, zero_fill))
    return out.reshape(shape)


@_stencil_call.defjvp
def _stencil_call_jvp(derivative, zero_fill, primals, tangents):
    ip, qs = primals
    _, dqs = tangents
    out = _stencil_call(ip, qs, derivative, zero_fill)
    tan = jnp.zeros_like(out)
    for ax, dq in enumerate(dqs):
        if type(dq) is jax.custom_derivatives.SymbolicZero:
            continue
        d1 = tuple(d + (1 if k == ax else 0) for k, d in enumerate(derivative))
        tan = tan + _stencil_call(ip, qs, d1, True) * jnp.broadcast_to(dq, out.shape)  # fills: zero tangent
    return out, tan


class Interpolator1D(_InterpolatorND):
    """interpax.Interpolator1D: Interpolator1D(x, f, method, extrap, period, **kwargs)(xq, dx=0)."""

    _ndim = 1

    def __init__(self, x, f, method="cubic", extrap=False, period=None, **kwargs):
        super().__init__(x, f, method=method, extrap=extrap, period=period, **kwargs)


class Interpolator2D(_InterpolatorND):
    """interpax.Interpolator2D: (x, y, f, ...)(xq, yq, dx=0, dy=0)."""

    _ndim = 2

    def __init__(self, x, y, f, method="cubic", extrap=False, period=None, **kwargs):
        super().__init__(x, y, f, method=method, extrap=extrap, period=period, **kwargs)


class Interpolator3D(_InterpolatorND):
    """interpax.Interpolator3D: (x, y, z, f, ...)(xq, yq, zq, dx=0, dy=0, dz=0)."""

    _ndim = 3

    def __init__(self, x, y, z, f, method="cubic", extrap=False, period=None, **kwargs):
        super().__init__(x, y, z, f, method=method, extrap=extrap, period=period, **kwargs)


for _cls in (Interpolator1D, Interpolator2D, Interpolator3D):
    jax.tree_util.register_pytree_node_class(_cls)

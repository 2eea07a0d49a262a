"""NumPy restatement of interpax's query-evaluation path.  TEST INFRASTRUCTURE ONLY.

This module is the parity checker for the CUDA path. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it. Nothing under ``interpax_b200/`` imports it, and the product
path has no CPU fallback.

It follows the reference's *operation order* (reciprocal-then-multiply for the
local coordinate, derivative tables scaled by the cell width before the
contraction, dense ``A @ F`` followed by a monomial einsum), so it is the closest
proxy for what XLA computes that can exist in an environment without JAX.

Pinning status (see DESIGN.md "Oracle"):
  * the reference itself cannot be imported here (jax, jaxlib, equinox, lineax are
    absent and there is no network), so values below ~1e-6 are pinned only by this
    restatement: **parity unpinned below the reference tests' own tolerances**;
  * pinned against: the reference's golden vectors (Akima table, PCHIP NAG table,
    README example and every analytic-function case of tests/test_interpolate.py
    with the reference's tolerances), ``A_CUBIC/A_BICUBIC/A_TRICUBIC`` compared
    element-wise with ``/root/reference/interpax/_coefs.py`` (numpy-only, importable),
    SciPy 1.18 ``CubicSpline / PchipInterpolator / Akima1DInterpolator /
    CubicHermiteSpline`` for slopes and 1-D values, ``np.searchsorted`` for indices.

Reference citations are ``/root/reference/interpax/<file>:<lines>``.
"""

from __future__ import annotations

import numpy as np

CUBIC_METHODS = (
    "cubic",
    "cubic2",
    "cardinal",
    "catmull-rom",
    "akima",
    "monotonic",
    "monotonic-0",
)
OTHER_METHODS = ("nearest", "linear")
METHODS = CUBIC_METHODS + OTHER_METHODS

# ---------------------------------------------------------------------------------
# _coefs.py:8-163 -- the dense coefficient matrices.  A_CUBIC is written out; the
# 16x16 and 64x64 ones are generated as tensor powers under the reference's index
# order and compared element-wise with the reference file in
# tests/test_oracle_pins.py (they are identical).
# ---------------------------------------------------------------------------------
A_CUBIC = np.array(
    [[1, 0, 0, 0], [0, 0, 1, 0], [-3, 3, -2, -1], [2, -2, 1, 1]], dtype=np.int64
)

# derivative-table order used by the reference when it stacks F
#   2-D: f, fx, fy, fxy                      (_spline.py:779-783)
#   3-D: f, fx, fy, fz, fxy, fxz, fyz, fxyz  (_spline.py:1070-1078)
TABLES_2D = ("f", "fx", "fy", "fxy")
TABLES_3D = ("f", "fx", "fy", "fz", "fxy", "fxz", "fyz", "fxyz")


def _is_deriv(name: str, letter: str) -> int:
    return 1 if letter in name[1:] else 0


def _build_a_bicubic() -> np.ndarray:
    A = np.zeros((16, 16), dtype=np.int64)
    for ff, name in enumerate(TABLES_2D):
        a, b = _is_deriv(name, "x"), _is_deriv(name, "y")
        for jj in (0, 1):
            for ii in (0, 1):
                n = 4 * ff + 2 * jj + ii
                for q in range(4):
                    for p in range(4):
                        A[p + 4 * q, n] = A_CUBIC[p, 2 * a + ii] * A_CUBIC[q, 2 * b + jj]
    return A


def _build_a_tricubic() -> np.ndarray:
    A = np.zeros((64, 64), dtype=np.int64)
    for ff, name in enumerate(TABLES_3D):
        a, b, c = (_is_deriv(name, s) for s in "xyz")
        for kk in (0, 1):
            for jj in (0, 1):
                for ii in (0, 1):
                    n = 8 * ff + 4 * kk + 2 * jj + ii
                    for r in range(4):
                        for q in range(4):
                            for p in range(4):
                                A[p + 4 * q + 16 * r, n] = (
                                    A_CUBIC[p, 2 * a + ii]
                                    * A_CUBIC[q, 2 * b + jj]
                                    * A_CUBIC[r, 2 * c + kk]
                                )
    return A


A_BICUBIC = _build_a_bicubic()
A_TRICUBIC = _build_a_tricubic()


# ---------------------------------------------------------------------------------
# utils.py:17-50
# ---------------------------------------------------------------------------------
def isbool(x) -> bool:
    """utils.py:17-19."""
    return isinstance(x, bool) or (hasattr(x, "dtype") and (x.dtype == bool))


def asarray_inexact(x):
    """utils.py:42-50.  Integer/bool arrays become the default float (f64 with x64)."""
    x = np.asarray(x)
    if not np.issubdtype(x.dtype, np.inexact):
        x = x.astype(np.result_type(x.dtype, np.float64))
    return x


def _real_dtype(dt):
    return np.zeros((), dtype=dt).real.dtype


# ---------------------------------------------------------------------------------
# _spline.py:1138-1222 helpers
# ---------------------------------------------------------------------------------
def _get_t_der(t, derivative: int, dxi):
    """_spline.py:1138-1151: rows [1,t,t^2,t^3] or their n-th derivative times dxi**n.

    ``lax.switch`` clamps the branch index, so order >= 4 gives zeros and a negative
    order gives branch 0.
    """
    t0 = np.zeros_like(t)
    t1 = np.ones_like(t)
    dxi = np.atleast_1d(dxi)[:, None]
    n = int(min(max(int(derivative), 0), 4))
    if n == 0:
        return np.array([t1, t, t**2, t**3]).T * dxi**0
    if n == 1:
        return np.array([t0, t1, 2 * t, 3 * t**2]).T * dxi
    if n == 2:
        return np.array([t0, t0, 2 * t1, 6 * t]).T * dxi**2
    if n == 3:
        return np.array([t0, t0, t0, 6 * t1]).T * dxi**3
    return np.array([t0, t0, t0, t0]).T * (dxi * 0)


def _parse_ndarg(arg, n):
    """_spline.py:1154-1161."""
    try:
        k = len(arg)
    except TypeError:
        arg = tuple(arg for _ in range(n))
        k = n
    assert k == n, "got too many args"
    return arg


def _parse_extrap(extrap, n):
    """_spline.py:1164-1177."""
    if isbool(extrap):
        return tuple(extrap for _ in range(2 * n))
    elif np.isscalar(extrap):
        return tuple(extrap for _ in range(2 * n))
    elif len(extrap) == 2 and np.isscalar(extrap[0]):
        return tuple(e for _ in range(n) for e in extrap)
    elif len(extrap) == n and all(len(extrap[i]) == 2 for i in range(n)):
        return tuple(eij for ei in extrap for eij in ei)
    raise ValueError(
        "extrap should either be a scalar, 2 element sequence (lo, hi), "
        + "or a sequence with 2 elements for each dimension"
    )


def _extrap(xq, fq, x, lo, hi):
    """_spline.py:1180-1222: post-mask of out-of-range queries along one axis."""
    if not (isbool(lo) and bool(lo)):
        val = np.nan if isbool(lo) else lo
        fq = np.where(xq < x[0], val, fq.T).T
    if not (isbool(hi) and bool(hi)):
        val = np.nan if isbool(hi) else hi
        fq = np.where(xq > x[-1], val, fq.T).T
    return fq


def _make_periodic(xq, x, period, axis, *arrs):
    """_spline.py:1108-1135: wrap queries and knots, sort knots, wrap-pad every table."""
    period = abs(period)
    dt = x.dtype
    xq = (xq % np.asarray(period, dtype=dt)).astype(dt)
    x = (x % np.asarray(period, dtype=dt)).astype(dt)
    i = np.argsort(x, kind="stable")
    x = x[i]
    x = np.concatenate([x[-1:] - np.asarray(period, dt), x, x[:1] + np.asarray(period, dt)])
    out = []
    for a in arrs:
        if a is not None:
            a = np.take(a, i, axis, mode="wrap")
            a = np.concatenate(
                [np.take(a, np.array([-1]), axis), a, np.take(a, np.array([0]), axis)],
                axis=axis,
            )
        out.append(a)
    return (xq, x, *out)


def bin_index(x, q):
    """a1: ``clip(searchsorted(x, q, side="right"), 1, N-1)`` (_spline.py:543,571,...)."""
    return np.clip(np.searchsorted(x, q, side="right"), 1, len(x) - 1)


def _recip(dx):
    """``where(dx == 0, 0, 1/dx)`` (_spline.py:546, 578, 741, 1057, ...)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(dx == 0, 0, 1 / dx).astype(dx.dtype)


def _rowscale(a, s):
    """``(a.T * s).T``: multiply along the leading (query) axis."""
    return (a.T * s).T


def _front(xs, fs):
    """dtype front matter shared by interp1d/2d/3d (_spline.py:503-506, 663-666, 882-885)."""
    # Python scalars are weakly typed in JAX (utils.py:45-46) exactly as they are for
    # np.result_type under NumPy 2 promotion, so hand them over un-converted.
    def item(a):
        if isinstance(a, (bool, int, float, complex)):
            return a
        return asarray_inexact(a)

    its = [item(a) for a in xs]
    xtype = np.result_type(*its)
    if not np.issubdtype(xtype, np.inexact):
        xtype = np.dtype(np.float64)
    xs = [np.asarray(a).astype(xtype) for a in its]
    fi = item(fs)
    ftype = np.result_type(fi, xtype)
    f = np.asarray(fi).astype(ftype)
    return xs, f


# ---------------------------------------------------------------------------------
# interp1d   _spline.py:444-592
# ---------------------------------------------------------------------------------
def interp1d(xq, x, f, method="cubic", derivative=0, extrap=False, period=None, **kwargs):
    (xq, x), f = _front((xq, x), f)
    axis = kwargs.get("axis", 0)
    fx = kwargs.pop("fx", None)
    outshape = xq.shape + f.shape[1:]
    xq = np.atleast_1d(xq)
    if (len(x) != f.shape[axis]) or (np.ndim(x) != 1):
        raise ValueError("x and f must be arrays of equal length")
    if method not in METHODS:
        raise ValueError(f"unknown method {method}")

    lowx, highx = _parse_extrap(extrap, 1)
    if period is not None:
        xq, x, f, fx = _make_periodic(xq, x, period, axis, f, fx)
        lowx = highx = True

    derivative = int(derivative)
    if method == "nearest":
        if max(derivative, 0) == 0:
            # _spline.py:531-533: brute-force argmin over all knots, first minimum wins
            i = np.argmin(np.abs(xq[:, None] - x[None]), axis=1)
            fq = f[i]
        else:
            fq = np.zeros((xq.size, *f.shape[1:]), dtype=f.dtype)
    elif method == "linear":
        n = min(max(derivative, 0), 2)
        if n == 2:
            fq = np.zeros((xq.size, *f.shape[1:]), dtype=f.dtype)
        else:
            i = bin_index(x, xq)
            df = np.take(f, i, axis) - np.take(f, i - 1, axis)
            dx = x[i] - x[i - 1]
            dxi = _recip(dx)
            if n == 0:
                delta = xq - x[i - 1]
                fq = np.where(
                    (dx == 0),
                    np.take(f, i, axis).T,
                    np.take(f, i - 1, axis).T + (delta * dxi * df.T),
                ).T
            else:
                fq = (df.T * dxi).T
    else:
        i = bin_index(x, xq)
        if fx is None:
            fx = approx_df(x, f, method, axis, **kwargs)
        assert fx.shape == f.shape
        dx = x[i] - x[i - 1]
        delta = xq - x[i - 1]
        dxi = _recip(dx)
        t = delta * dxi
        f0 = np.take(f, i - 1, axis)
        f1 = np.take(f, i, axis)
        fx0 = _rowscale(np.take(fx, i - 1, axis), dx)
        fx1 = _rowscale(np.take(fx, i, axis), dx)
        F = np.stack([f0, f1, fx0, fx1], axis=0)  # (4, Nq, *batch)
        coef = np.tensordot(A_CUBIC.astype(F.dtype), F, axes=(1, 0))  # (4, Nq, *batch)
        ttx = _get_t_der(t, derivative, dxi)
        fq = np.einsum("ji...,ij->i...", coef, ttx)

    fq = _extrap(xq, fq, x, lowx, highx)
    return fq.reshape(outshape)


# ---------------------------------------------------------------------------------
# interp2d   _spline.py:595-805
# ---------------------------------------------------------------------------------
def interp2d(xq, yq, x, y, f, method="cubic", derivative=0, extrap=False, period=None, **kwargs):
    (xq, yq, x, y), f = _front((xq, yq, x, y), f)
    xtype = x.dtype
    fx = kwargs.pop("fx", None)
    fy = kwargs.pop("fy", None)
    fxy = kwargs.pop("fxy", None)
    xq, yq = np.broadcast_arrays(xq, yq)
    outshape = xq.shape + f.shape[2:]
    xq, yq = map(np.atleast_1d, (xq, yq))
    if (len(x) != f.shape[0]) or (x.ndim != 1):
        raise ValueError("x and f must be arrays of equal length")
    if (len(y) != f.shape[1]) or (y.ndim != 1):
        raise ValueError("y and f must be arrays of equal length")
    if method not in METHODS:
        raise ValueError(f"unknown method {method}")

    periodx, periody = _parse_ndarg(period, 2)
    derivative_x, derivative_y = (int(d) for d in _parse_ndarg(derivative, 2))
    lowx, highx, lowy, highy = _parse_extrap(extrap, 2)

    if periodx is not None:
        xq, x, f, fx, fy, fxy = _make_periodic(xq, x, periodx, 0, f, fx, fy, fxy)
        lowx = highx = True
    if periody is not None:
        yq, y, f, fx, fy, fxy = _make_periodic(yq, y, periody, 1, f, fx, fy, fxy)
        lowy = highy = True

    if method == "nearest":
        if derivative_x == 0 and derivative_y == 0:
            # _spline.py:704-719: nearest of the 4 cell corners, first minimum wins
            i = bin_index(x, xq)
            j = bin_index(y, yq)
            nbx = np.array([[x[i], x[i - 1], x[i], x[i - 1]], [y[j], y[j], y[j - 1], y[j - 1]]])
            nbf = [f[i, j], f[i - 1, j], f[i, j - 1], f[i - 1, j - 1]]
            xyq = np.array([xq, yq])
            d = nbx - xyq[:, None, :]
            dist = np.sqrt(np.sum(np.abs(d) ** 2, axis=0))
            idx = np.argmin(dist, axis=0)
            fq = np.zeros((xq.size, *f.shape[2:]), dtype=f.dtype)
            for c in range(4):
                m = idx == c
                fq[m] = nbf[c][m]
        else:
            fq = np.zeros((xq.size, *f.shape[2:]), dtype=f.dtype)
    elif method == "linear":
        i = bin_index(x, xq)
        j = bin_index(y, yq)
        f00 = f[i - 1, j - 1]
        f01 = f[i - 1, j]
        f10 = f[i, j - 1]
        f11 = f[i, j]
        x0, x1, y0, y1 = x[i - 1], x[i], y[j - 1], y[j]
        dx = x1 - x0
        dxi = _recip(dx)
        dy = y1 - y0
        dyi = _recip(dy)

        def lin_w(n, lo, hi, q):
            # _spline.py:745-753 (lax.switch clamps to [0, 2])
            n = min(max(n, 0), 2)
            if n == 0:
                return np.array([hi - q, q - lo])
            if n == 1:
                return np.array([-np.ones_like(q), np.ones_like(q)])
            return np.zeros((2, q.size), dtype=xtype)

        tx = lin_w(derivative_x, x0, x1, xq)
        ty = lin_w(derivative_y, y0, y1, yq)
        F = np.array([[f00, f01], [f10, f11]])
        fq = (dxi * dyi * np.einsum("ijk...,ik,jk->k...", F, tx, ty).T).T
    else:
        if fx is None:
            fx = approx_df(x, f, method, 0, **kwargs)
        if fy is None:
            fy = approx_df(y, f, method, 1, **kwargs)
        if fxy is None:
            fxy = approx_df(y, fx, method, 1, **kwargs)
        assert fx.shape == fy.shape == fxy.shape == f.shape
        i = bin_index(x, xq)
        j = bin_index(y, yq)
        dx = x[i] - x[i - 1]
        deltax = xq - x[i - 1]
        dxi = _recip(dx)
        tx = deltax * dxi
        dy = y[j] - y[j - 1]
        deltay = yq - y[j - 1]
        dyi = _recip(dy)
        ty = deltay * dyi

        tabs = dict(zip(TABLES_2D, (f, fx, fy, fxy)))
        cols = []
        for name in TABLES_2D:
            for jj in (0, 1):
                for ii in (0, 1):
                    v = tabs[name][i - 1 + ii, j - 1 + jj]
                    if "x" in name[1:]:
                        v = _rowscale(v, dx)
                    if "y" in name[1:]:
                        v = _rowscale(v, dy)
                    cols.append(v)
        F = np.stack(cols, axis=0)  # (16, Nq, *batch)
        coef = np.tensordot(A_BICUBIC.astype(F.dtype), F, axes=(1, 0))  # (16, Nq, ...)
        # Fortran-order unpack: coef index p + 4 q  ->  [p, q]   (_spline.py:797)
        coef = coef.reshape((4, 4, *coef.shape[1:]), order="F")
        coef = np.moveaxis(coef, 2, 0)  # (Nq, 4, 4, *batch)
        ttx = _get_t_der(tx, derivative_x, dxi)
        tty = _get_t_der(ty, derivative_y, dyi)
        fq = np.einsum("ijk...,ij,ik->i...", coef, ttx, tty)

    fq = _extrap(xq, fq, x, lowx, highx)
    fq = _extrap(yq, fq, y, lowy, highy)
    return fq.reshape(outshape)


# ---------------------------------------------------------------------------------
# interp3d   _spline.py:808-1105
# ---------------------------------------------------------------------------------
def interp3d(
    xq, yq, zq, x, y, z, f, method="cubic", derivative=0, extrap=False, period=None, **kwargs
):
    (xq, yq, zq, x, y, z), f = _front((xq, yq, zq, x, y, z), f)
    xtype = x.dtype
    if (len(x) != f.shape[0]) or (x.ndim != 1):
        raise ValueError("x and f must be arrays of equal length")
    if (len(y) != f.shape[1]) or (y.ndim != 1):
        raise ValueError("y and f must be arrays of equal length")
    if (len(z) != f.shape[2]) or (z.ndim != 1):
        raise ValueError("z and f must be arrays of equal length")
    if method not in METHODS:
        raise ValueError(f"unknown method {method}")
    xq, yq, zq = np.broadcast_arrays(xq, yq, zq)
    outshape = xq.shape + f.shape[3:]
    xq, yq, zq = map(np.atleast_1d, (xq, yq, zq))

    tabs = {"f": f}
    for name in TABLES_3D[1:]:
        tabs[name] = kwargs.pop(name, None)

    periodx, periody, periodz = _parse_ndarg(period, 3)
    derivative_x, derivative_y, derivative_z = (int(d) for d in _parse_ndarg(derivative, 3))
    lowx, highx, lowy, highy, lowz, highz = _parse_extrap(extrap, 3)

    def wrap(q, knots, per, ax):
        out = _make_periodic(q, knots, per, ax, *[tabs[n] for n in TABLES_3D])
        for n, a in zip(TABLES_3D, out[2:]):
            tabs[n] = a
        return out[0], out[1]

    if periodx is not None:
        xq, x = wrap(xq, x, periodx, 0)
        lowx = highx = True
    if periody is not None:
        yq, y = wrap(yq, y, periody, 1)
        lowy = highy = True
    if periodz is not None:
        zq, z = wrap(zq, z, periodz, 2)
        lowz = highz = True
    f = tabs["f"]

    if method == "nearest":
        if derivative_x == 0 and derivative_y == 0 and derivative_z == 0:
            # _spline.py:942-971: nearest of the 8 cell corners, first minimum wins
            i = bin_index(x, xq)
            j = bin_index(y, yq)
            k = bin_index(z, zq)
            nbx = np.array(
                [
                    [x[i], x[i - 1], x[i], x[i - 1], x[i], x[i - 1], x[i], x[i - 1]],
                    [y[j], y[j], y[j - 1], y[j - 1], y[j], y[j], y[j - 1], y[j - 1]],
                    [z[k], z[k], z[k], z[k], z[k - 1], z[k - 1], z[k - 1], z[k - 1]],
                ]
            )
            nbf = [
                f[i, j, k],
                f[i - 1, j, k],
                f[i, j - 1, k],
                f[i - 1, j - 1, k],
                f[i, j, k - 1],
                f[i - 1, j, k - 1],
                f[i, j - 1, k - 1],
                f[i - 1, j - 1, k - 1],
            ]
            xyzq = np.array([xq, yq, zq])
            d = nbx - xyzq[:, None, :]
            dist = np.sqrt(np.sum(np.abs(d) ** 2, axis=0))
            idx = np.argmin(dist, axis=0)
            fq = np.zeros((xq.size, *f.shape[3:]), dtype=f.dtype)
            for c in range(8):
                m = idx == c
                fq[m] = nbf[c][m]
        else:
            fq = np.zeros((xq.size, *f.shape[3:]), dtype=f.dtype)
    elif method == "linear":
        i = bin_index(x, xq)
        j = bin_index(y, yq)
        k = bin_index(z, zq)
        f000 = f[i - 1, j - 1, k - 1]
        f001 = f[i - 1, j - 1, k]
        f010 = f[i - 1, j, k - 1]
        f100 = f[i, j - 1, k - 1]
        f110 = f[i, j, k - 1]
        f011 = f[i - 1, j, k]
        f101 = f[i, j - 1, k]
        f111 = f[i, j, k]
        x0, x1, y0, y1, z0, z1 = x[i - 1], x[i], y[j - 1], y[j], z[k - 1], z[k]
        dx = x1 - x0
        dxi = _recip(dx)
        dy = y1 - y0
        dyi = _recip(dy)
        dz = z1 - z0
        dzi = _recip(dz)

        def lin_w(n, lo, hi, q):
            n = min(max(n, 0), 2)
            if n == 0:
                return np.array([hi - q, q - lo])
            if n == 1:
                return np.array([-np.ones_like(q), np.ones_like(q)])
            return np.zeros((2, q.size), dtype=xtype)

        tx = lin_w(derivative_x, x0, x1, xq)
        ty = lin_w(derivative_y, y0, y1, yq)
        tz = lin_w(derivative_z, z0, z1, zq)
        F = np.array([[[f000, f001], [f010, f011]], [[f100, f101], [f110, f111]]])
        fq = (dxi * dyi * dzi * np.einsum("lijk...,lk,ik,jk->k...", F, tx, ty, tz).T).T
    else:
        # slope chain, _spline.py:1027-1040
        if tabs["fx"] is None:
            tabs["fx"] = approx_df(x, f, method, 0, **kwargs)
        if tabs["fy"] is None:
            tabs["fy"] = approx_df(y, f, method, 1, **kwargs)
        if tabs["fz"] is None:
            tabs["fz"] = approx_df(z, f, method, 2, **kwargs)
        if tabs["fxy"] is None:
            tabs["fxy"] = approx_df(y, tabs["fx"], method, 1, **kwargs)
        if tabs["fxz"] is None:
            tabs["fxz"] = approx_df(z, tabs["fx"], method, 2, **kwargs)
        if tabs["fyz"] is None:
            tabs["fyz"] = approx_df(z, tabs["fy"], method, 2, **kwargs)
        if tabs["fxyz"] is None:
            tabs["fxyz"] = approx_df(z, tabs["fxy"], method, 2, **kwargs)
        assert all(tabs[n].shape == f.shape for n in TABLES_3D)

        i = bin_index(x, xq)
        j = bin_index(y, yq)
        k = bin_index(z, zq)
        dx = x[i] - x[i - 1]
        deltax = xq - x[i - 1]
        dxi = _recip(dx)
        tx = deltax * dxi
        dy = y[j] - y[j - 1]
        deltay = yq - y[j - 1]
        dyi = _recip(dy)
        ty = deltay * dyi
        dz = z[k] - z[k - 1]
        deltaz = zq - z[k - 1]
        dzi = _recip(dz)
        tz = deltaz * dzi

        cols = []
        for name in TABLES_3D:
            for kk in (0, 1):
                for jj in (0, 1):
                    for ii in (0, 1):
                        v = tabs[name][i - 1 + ii, j - 1 + jj, k - 1 + kk]
                        if "x" in name[1:]:
                            v = _rowscale(v, dx)
                        if "y" in name[1:]:
                            v = _rowscale(v, dy)
                        if "z" in name[1:]:
                            v = _rowscale(v, dz)
                        cols.append(v)
        F = np.stack(cols, axis=0)  # (64, Nq, *batch)
        coef = np.tensordot(A_TRICUBIC.astype(F.dtype), F, axes=(1, 0))
        coef = coef.reshape((4, 4, 4, *coef.shape[1:]), order="F")
        coef = np.moveaxis(coef, 3, 0)  # (Nq, 4, 4, 4, *batch)
        ttx = _get_t_der(tx, derivative_x, dxi)
        tty = _get_t_der(ty, derivative_y, dyi)
        ttz = _get_t_der(tz, derivative_z, dzi)
        fq = np.einsum("lijk...,li,lj,lk->l...", coef, ttx, tty, ttz)

    fq = _extrap(xq, fq, x, lowx, highx)
    fq = _extrap(yq, fq, y, lowy, highy)
    fq = _extrap(zq, fq, z, lowz, highz)
    return fq.reshape(outshape)


# ---------------------------------------------------------------------------------
# approx_df   _fd_derivs.py:9-443
# ---------------------------------------------------------------------------------
def approx_df(x, f, method="cubic", axis=-1, **kwargs):
    """_fd_derivs.py:9-76."""
    x = np.asarray(x)
    f = np.asarray(f)
    if method == "cubic":
        return _cubic1(x, f, axis)
    if method == "cubic2":
        bc_type = kwargs.get("bc_type", "not-a-knot")
        ax = axis % f.ndim
        bc, dtype = _validate_bc(bc_type, f.shape[:ax] + f.shape[ax + 1 :], f.dtype)
        return _cubic2(x, f, axis, bc, dtype)
    if method == "cardinal":
        return _cardinal(x, f, axis, kwargs.get("c", 0))
    if method == "catmull-rom":
        return _cardinal(x, f, axis, 0)
    if method in ("monotonic", "monotonic-0"):
        zero = method == "monotonic-0"
        if np.iscomplexobj(f):
            return _monotonic(x, f.real, axis, zero) + 1j * _monotonic(x, f.imag, axis, zero)
        return _monotonic(x, f, axis, zero)
    if method == "akima":
        return _akima(x, f, axis)
    if method in ("nearest", "linear"):
        return np.zeros_like(f)
    raise ValueError(f"got unknown method {method}")


def _along(v, ndim):
    """Reshape a 1-D per-interval vector so it broadcasts along axis 0 of an ndim array."""
    return v.reshape((v.shape[0],) + (1,) * (ndim - 1))


def _cubic1(x, f, axis):
    """_fd_derivs.py:79-101: mean of the two adjacent secant slopes, one-sided ends."""
    f = np.moveaxis(f, axis, 0)
    dx = np.diff(x)
    df = np.diff(f, axis=0)
    dxi = _along(_recip(dx), f.ndim)
    df = dxi * df
    fx = np.concatenate([df[:1], 1 / 2 * (df[:-1] + df[1:]), df[-1:]], axis=0)
    return np.moveaxis(fx, 0, axis)


def _cardinal(x, f, axis, c=0.0):
    """_fd_derivs.py:303-324."""
    f = np.moveaxis(f, axis, 0)
    dx = x[2:] - x[:-2]
    df = f[2:] - f[:-2]
    dxi = _along(_recip(dx), f.ndim)
    df = dxi * df
    with np.errstate(divide="ignore", invalid="ignore"):
        fx0 = (f[1:2] - f[0:1]) * np.where(x[0] == x[1], 0, 1 / (x[1] - x[0]))
        fx1 = (f[-1:] - f[-2:-1]) * np.where(x[-1] == x[-2], 0, 1 / (x[-1] - x[-2]))
    fx = (1 - c) * np.concatenate([fx0, df, fx1], axis=0)
    return np.moveaxis(fx.astype(f.dtype), 0, axis)


def safediv(a, b, fill=0, threshold=0):
    """_fd_derivs.py:427-443."""
    mask = np.abs(b) <= threshold
    num = np.where(mask, fill, a)
    den = np.where(mask, 1, b)
    return num / den


def _monotonic(x, f, axis, zero_slope):
    """_fd_derivs.py:327-387 (PCHIP / Fritsch-Butland; ends after Moler's pchiptx)."""
    f = np.moveaxis(f, axis, 0)
    fshp = f.shape
    f = f.reshape(f.shape[0], -1)
    x = x[:, None]
    hk = x[1:] - x[:-1]
    df = np.diff(f, axis=0)
    hki = _recip(hk)
    mk = hki * df
    smk = np.sign(mk)
    condition = (smk[1:, :] != smk[:-1, :]) | (mk[1:, :] == 0) | (mk[:-1, :] == 0)
    w1 = 2 * hk[1:] + hk[:-1]
    w2 = hk[1:] + 2 * hk[:-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        whmean = safediv(safediv(w1, mk[:-1, :]) + safediv(w2, mk[1:, :]), w1 + w2)
        dk = np.where(condition, 0, safediv(np.array(1.0, dtype=f.dtype), whmean))
    dk = dk.astype(f.dtype)
    if zero_slope:
        d0 = np.zeros((1, dk.shape[1]), dtype=f.dtype)
        d1 = np.zeros((1, dk.shape[1]), dtype=f.dtype)
    else:

        def edge(h0, h1, m0, m1):
            with np.errstate(divide="ignore", invalid="ignore"):
                d = ((2 * h0 + h1) * m0 - h0 * m1) / (h0 + h1)
            mask = np.sign(d) != np.sign(m0)
            mask2 = (np.sign(m0) != np.sign(m1)) & (np.abs(d) > 3.0 * np.abs(m0))
            mmm = (~mask) & mask2
            d = np.where(mask, 0.0, d)
            d = np.where(mmm, 3.0 * m0, d)
            return d

        with np.errstate(divide="ignore"):
            hk2 = 1 / hki
        n1 = hk2.shape[0]
        # JAX clamps out-of-range static indices on reads, which is what makes the
        # two-knot case work (_fd_derivs.py:379-380; pinned by test_scipy.py:710-717)
        ix = lambda k: min(max(k if k >= 0 else n1 + k, 0), n1 - 1)
        d0 = edge(hk2[ix(0)], hk2[ix(1)], mk[ix(0)], mk[ix(1)])[None]
        d1 = edge(hk2[ix(-1)], hk2[ix(-2)], mk[ix(-1)], mk[ix(-2)])[None]
    dk = np.concatenate([d0.astype(f.dtype), dk, d1.astype(f.dtype)])
    dk = dk.reshape(fshp)
    return np.moveaxis(dk, 0, axis)


def _akima(x, f, axis):
    """_fd_derivs.py:390-424.  ``jnp.empty`` is zero-filled; the threshold uses the
    maximum over the whole array (batch columns included)."""
    dx = np.diff(x)
    f = np.moveaxis(f, axis, 0)
    m = np.zeros((x.size + 3,) + f.shape[1:], dtype=f.dtype)
    dx = _along(dx, f.ndim)
    mask = dx == 0
    dx = np.where(mask, 1, dx)
    dxi = np.where(mask, 0.0, 1 / dx)
    m[2:-2] = np.diff(f, axis=0) * dxi
    m[1] = 2.0 * m[2] - m[3]
    m[0] = 2.0 * m[1] - m[2]
    m[-2] = 2.0 * m[-3] - m[-4]
    m[-1] = 2.0 * m[-2] - m[-3]
    dm = np.abs(np.diff(m, axis=0))
    m2 = m[1:-2]
    m3 = m[2:-1]
    m4m3 = dm[2:]
    m2m1 = dm[:-2]
    f12 = m4m3 + m2m1
    mask = f12 > 1e-9 * np.max(f12, initial=-np.inf)
    df = (m4m3 * m2 + m2m1 * m3) / np.where(mask, f12, 1.0)
    df = np.where(mask, df, 0.5 * (m[3:] + m[:-3]))
    return np.moveaxis(df.astype(f.dtype), 0, axis)


def _validate_bc(bc_type, expected_deriv_shape, dtype):
    """_fd_derivs.py:104-157."""
    if isinstance(bc_type, str):
        if bc_type == "periodic":
            raise NotImplementedError
        bc_type = (bc_type, bc_type)
    else:
        if len(bc_type) != 2:
            raise ValueError(
                "`bc_type` must contain 2 elements to specify start and end conditions."
            )
        if "periodic" in bc_type:
            raise ValueError(
                "'periodic' `bc_type` is defined for both "
                + "curve ends and cannot be used with other "
                + "boundary conditions."
            )
    validated = []
    for bc in bc_type:
        if isinstance(bc, str):
            if bc == "clamped":
                validated.append((1, np.zeros(expected_deriv_shape)))
            elif bc == "natural":
                validated.append((2, np.zeros(expected_deriv_shape)))
            elif bc in ["not-a-knot", "periodic"]:
                validated.append(bc)
            else:
                raise ValueError(f"bc_type={bc} is not allowed.")
        else:
            try:
                deriv_order, deriv_value = bc
            except Exception as e:
                raise ValueError(
                    "A specified derivative value must be given in the form (order, value)."
                ) from e
            if deriv_order not in [1, 2]:
                raise ValueError("The specified derivative order must be 1 or 2.")
            deriv_value = asarray_inexact(deriv_value)
            dtype = np.promote_types(dtype, deriv_value.dtype)
            if deriv_value.shape != expected_deriv_shape:
                raise ValueError(
                    "`deriv_value` shape {} is not the expected one {}.".format(
                        deriv_value.shape, expected_deriv_shape
                    )
                )
            validated.append((deriv_order, deriv_value))
    return validated, dtype


def thomas(diag, lower, upper, b):
    """Tridiagonal solve without pivoting, in the order lineax's ``Tridiagonal`` solver
    uses (lineax 0.0.5-0.1.0, third-party, not in /root/reference; called at
    _fd_derivs.py:295-298): forward sweep ``den = b_i - a_i c'_{i-1}``,
    ``d'_i = (d_i - a_i d'_{i-1}) / den``, ``c'_i = c_i / den``; back substitution
    ``x_i = d'_i - c'_i x_{i+1}``.  ``b`` may carry trailing columns."""
    n = diag.shape[0]
    cp = np.zeros(n, dtype=b.dtype)
    dp = np.zeros_like(b)
    c_prev = 0
    d_prev = np.zeros(b.shape[1:], dtype=b.dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(n):
            a = lower[k - 1] if k > 0 else 0
            c = upper[k] if k < n - 1 else 0
            den = diag[k] - a * c_prev
            d_prev = (b[k] - a * d_prev) / den
            c_prev = c / den
            cp[k] = c_prev
            dp[k] = d_prev
    xs = np.zeros_like(b)
    x_next = np.zeros(b.shape[1:], dtype=b.dtype)
    for k in range(n - 1, -1, -1):
        x_next = dp[k] - cp[k] * x_next
        xs[k] = x_next
    return xs


def _cubic2(x, f, axis, bc, dtype):
    """_fd_derivs.py:160-300: C2 spline slopes from the tridiagonal system."""
    fdtype = f.dtype
    f = f.astype(dtype)
    f = np.moveaxis(f, axis, 0)
    dx = np.diff(x)
    df = np.diff(f, axis=0)
    dxr = dx.reshape([dx.shape[0]] + [1] * (f.ndim - 1))
    with np.errstate(divide="ignore", invalid="ignore"):
        dxi = np.where(dxr == 0, 0, 1 / np.where(dxr == 0, 1, dxr))
    df = dxi * df
    n = len(f)
    bc = list(bc)
    if n == 2:
        if isinstance(bc[0], str) and bc[0] in ["not-a-knot", "periodic"]:
            bc[0] = (1, df[0])
        if isinstance(bc[1], str) and bc[1] in ["not-a-knot", "periodic"]:
            bc[1] = (1, df[0])

    if n == 3 and isinstance(bc[0], str) and isinstance(bc[1], str):
        # both not-a-knot with three knots: the parabola through the points
        A = np.zeros((3, 3), dtype=x.dtype)
        A[0, 0] = 1
        A[0, 1] = 1
        A[1, 0] = dx[1]
        A[1, 1] = 2 * (dx[0] + dx[1])
        A[1, 2] = dx[0]
        A[2, 1] = 1
        A[2, 2] = 1
        b = np.zeros((3,) + f.shape[1:], dtype=dtype)
        b[0] = 2 * df[0]
        b[1] = 3 * (dxr[0] * df[1] + dxr[1] * df[0])
        b[2] = 2 * df[1]
        fx = np.linalg.solve(A, b.reshape(3, -1)).reshape(b.shape)
        return np.moveaxis(fx, 0, axis).astype(fdtype)

    diag = np.zeros(n, dtype=x.dtype)
    diag[1:-1] = 2 * (dx[:-1] + dx[1:])
    upper = np.zeros(n - 1, dtype=x.dtype)
    upper[1:] = dx[:-1]
    lower = np.zeros(n - 1, dtype=x.dtype)
    lower[:-1] = dx[1:]
    b = np.zeros((n,) + f.shape[1:], dtype=dtype)
    b[1:-1] = 3 * (dxr[1:] * df[:-1] + dxr[:-1] * df[1:])

    bc_start, bc_end = bc
    if isinstance(bc_start, str):  # not-a-knot
        d = x[2] - x[0]
        diag[0] = dx[1]
        upper[0] = d
        b[0] = ((dxr[0] + 2 * d) * dxr[1] * df[0] + dxr[0] ** 2 * df[1]) / d
    elif bc_start[0] == 1:
        diag[0] = 1
        upper[0] = 0
        b[0] = bc_start[1]
    elif bc_start[0] == 2:
        diag[0] = 2 * dx[0]
        upper[0] = dx[0]
        b[0] = -0.5 * bc_start[1] * dx[0] ** 2 + 3 * (f[1] - f[0])

    if isinstance(bc_end, str):
        d = x[-1] - x[-3]
        diag[-1] = dx[-2]
        lower[-1] = d
        b[-1] = (dxr[-1] ** 2 * df[-2] + (2 * d + dxr[-1]) * dxr[-2] * df[-1]) / d
    elif bc_end[0] == 1:
        diag[-1] = 1
        lower[-1] = 0
        b[-1] = bc_end[1]
    elif bc_end[0] == 2:
        diag[-1] = 2 * dx[-1]
        lower[-1] = dx[-1]
        b[-1] = 0.5 * bc_end[1] * dx[-1] ** 2 + 3 * (f[-1] - f[-2])

    # duplicate knots: neutralise singular rows (_fd_derivs.py:281-286)
    mask = diag == 0
    diag = np.where(mask, 1, diag)
    lower = np.where(mask[1:], 0, lower)
    upper = np.where(mask[:-1], 0, upper)
    b = np.where(mask, 0, b.T).T

    rdtype = np.result_type(diag, lower, upper, b)
    fx = thomas(diag.astype(rdtype), lower.astype(rdtype), upper.astype(rdtype), b.astype(rdtype))
    return np.moveaxis(fx, 0, axis).astype(fdtype)


# ---------------------------------------------------------------------------------
# Interpolator1D/2D/3D   _spline.py:48-441  (thin containers; slopes precomputed on
# the *un-padded* tables, so a periodic class call differs from the functional form
# in the boundary cells -- see _spline.py:380-393 vs :924-938 + :1027-1040)
# ---------------------------------------------------------------------------------
class Interpolator1D:
    def __init__(self, x, f, method="cubic", extrap=False, period=None, **kwargs):
        x, f = map(asarray_inexact, (x, f))
        axis = kwargs.get("axis", 0)
        fx = kwargs.pop("fx", None)
        if (len(x) != f.shape[axis]) or (np.ndim(x) != 1):
            raise ValueError("x and f must be arrays of equal length")
        if method not in METHODS:
            raise ValueError(f"unknown method {method}")
        self.x, self.f, self.axis, self.method = x, f, axis, method
        self.extrap, self.period = extrap, period
        if fx is None:
            fx = approx_df(x, f, method, axis, **kwargs)
        self.derivs = {"fx": fx}

    def __call__(self, xq, dx=0):
        return interp1d(
            xq, self.x, self.f, self.method, dx, self.extrap, self.period, **self.derivs
        )


class Interpolator2D:
    def __init__(self, x, y, f, method="cubic", extrap=False, period=None, **kwargs):
        x, y, f = map(asarray_inexact, (x, y, f))
        fx = kwargs.pop("fx", None)
        fy = kwargs.pop("fy", None)
        fxy = kwargs.pop("fxy", None)
        if (len(x) != f.shape[0]) or (x.ndim != 1):
            raise ValueError("x and f must be arrays of equal length")
        if (len(y) != f.shape[1]) or (y.ndim != 1):
            raise ValueError("y and f must be arrays of equal length")
        if method not in METHODS:
            raise ValueError(f"unknown method {method}")
        self.x, self.y, self.f, self.method = x, y, f, method
        self.extrap, self.period = extrap, period
        if fx is None:
            fx = approx_df(x, f, method, 0, **kwargs)
        if fy is None:
            fy = approx_df(y, f, method, 1, **kwargs)
        if fxy is None:
            fxy = approx_df(y, fx, method, 1, **kwargs)
        self.derivs = {"fx": fx, "fy": fy, "fxy": fxy}

    def __call__(self, xq, yq, dx=0, dy=0):
        return interp2d(
            xq, yq, self.x, self.y, self.f, self.method, (dx, dy), self.extrap, self.period,
            **self.derivs,
        )


class Interpolator3D:
    def __init__(self, x, y, z, f, method="cubic", extrap=False, period=None, **kwargs):
        x, y, z, f = map(asarray_inexact, (x, y, z, f))
        if (len(x) != f.shape[0]) or (x.ndim != 1):
            raise ValueError("x and f must be arrays of equal length")
        if (len(y) != f.shape[1]) or (y.ndim != 1):
            raise ValueError("y and f must be arrays of equal length")
        if (len(z) != f.shape[2]) or (z.ndim != 1):
            raise ValueError("z and f must be arrays of equal length")
        if method not in METHODS:
            raise ValueError(f"unknown method {method}")
        d = {n: kwargs.pop(n, None) for n in TABLES_3D[1:]}
        self.x, self.y, self.z, self.f, self.method = x, y, z, f, method
        self.extrap, self.period = extrap, period
        if d["fx"] is None:
            d["fx"] = approx_df(x, f, method, 0, **kwargs)
        if d["fy"] is None:
            d["fy"] = approx_df(y, f, method, 1, **kwargs)
        if d["fz"] is None:
            d["fz"] = approx_df(z, f, method, 2, **kwargs)
        if d["fxy"] is None:
            d["fxy"] = approx_df(y, d["fx"], method, 1, **kwargs)
        if d["fxz"] is None:
            d["fxz"] = approx_df(z, d["fx"], method, 2, **kwargs)
        if d["fyz"] is None:
            d["fyz"] = approx_df(z, d["fy"], method, 2, **kwargs)
        if d["fxyz"] is None:
            d["fxyz"] = approx_df(z, d["fxy"], method, 2, **kwargs)
        self.derivs = d

    def __call__(self, xq, yq, zq, dx=0, dy=0, dz=0):
        return interp3d(
            xq, yq, zq, self.x, self.y, self.z, self.f, self.method, (dx, dy, dz),
            self.extrap, self.period, **self.derivs,
        )

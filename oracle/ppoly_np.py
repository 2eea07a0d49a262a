"""NumPy restatement of interpax's piecewise-polynomial family.  TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/interpax/_ppoly.py`` (PPoly :50-441, prepare_input :444-499,
CubicHermiteSpline :502-569, PchipInterpolator :572-652, Akima1DInterpolator :655-710,
CubicSpline :713-809) operation by operation: ``searchsorted(side="right")`` + clip,
``t = x - x[i-1]``, ``polyder`` (coefficient times falling factorial, one rounding),
``polyval`` (Horner, multiply then add), NaN outside unless extrapolating, the periodic
remainder ``x0 + (x - x0) % (xe - x0)``.

Only ``tests/`` import this.  Pinning: the reference's API and docstrings are SciPy's, and
the reference's own tests (tests/test_scipy.py) check it against SciPy's classes and their
known-answer tables; ``tests/test_oracle_ppoly.py`` pins this restatement against SciPy 1.18
``PPoly / CubicHermiteSpline / PchipInterpolator / Akima1DInterpolator / CubicSpline`` on the
same inputs (values, derivatives, antiderivatives, definite integrals).  The reference itself
cannot be imported here (no jax), so differences below those tests' tolerances are unpinned.
"""

from __future__ import annotations

import numpy as np

from .interpax_np import approx_df, asarray_inexact

A_CUBIC = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [-3, 3, -2, -1], [2, -2, 1, 1]])  # _coefs.py:8-15


def _real(dt):
    return np.float32 if np.dtype(dt) in (np.float32, np.complex64) else np.float64


def polyder_coeffs(k: int, nu: int) -> np.ndarray:
    """Factors jnp.polyder multiplies the leading k - nu coefficients with
    (highest power first): p (p-1) ... (p-nu+1) for p = k-1 .. nu."""
    p = np.arange(nu, k)[::-1, None] - np.arange(nu)[None, :]
    return p.prod(axis=1) if nu > 0 else np.ones(k, dtype=np.int64)


class PPoly:
    """_ppoly.py:50-441."""

    def __init__(self, c, x, extrapolate=None, axis=0, check=True):
        c = asarray_inexact(c)
        x = asarray_inexact(x)
        if c.ndim < 2:
            raise ValueError("Coefficients array must be at least 2-dimensional.")
        if x.ndim != 1:
            raise ValueError("x must be 1-dimensional")
        if x.size < 2:
            raise ValueError("at least 2 breakpoints are needed")
        axis = axis % (c.ndim - 1)
        if extrapolate is None:
            extrapolate = True
        elif extrapolate != "periodic":
            extrapolate = bool(extrapolate)
        if axis != 0:
            c = np.moveaxis(c, axis + 1, 0)
            c = np.moveaxis(c, axis + 1, 0)
        if c.shape[0] == 0:
            raise ValueError("polynomial must be at least of order 0")
        if c.shape[1] != x.size - 1:
            raise ValueError("number of coefficients != len(x)-1")
        if check and np.any(np.diff(x) < 0):
            raise ValueError("`x` must be strictly increasing sequence.")
        self.extrapolate = extrapolate
        self.axis = axis
        self.x = x
        self.c = c

    @classmethod
    def construct_fast(cls, c, x, extrapolate=None, axis=0):
        self = object.__new__(cls)
        self.c, self.x, self.axis = c, x, axis
        self.extrapolate = True if extrapolate is None else extrapolate
        return self

    def __call__(self, x, nu=0, extrapolate=None):
        if extrapolate is None:
            extrapolate = self.extrapolate
        x = asarray_inexact(x)
        dt = np.result_type(x.dtype, self.x.dtype)
        x = x.astype(dt)
        x_shape, x_ndim = x.shape, x.ndim
        x = x.ravel()
        if extrapolate == "periodic":
            x = self.x[0] + (x - self.x[0]) % (self.x[-1] - self.x[0])
            extrapolate = False
        i = np.clip(np.searchsorted(self.x, x, side="right"), 1, len(self.x) - 1)
        t = x - self.x[i - 1]
        c = self.c[:, i - 1]                      # (k, nq, ...)
        k = c.shape[0]
        out_dt = np.result_type(c.dtype, t.dtype)
        tt = t.reshape(t.shape + (1,) * (c.ndim - 2))
        y = np.zeros(c.shape[1:], dtype=out_dt)
        if nu < k:
            fac = polyder_coeffs(k, nu).astype(_real(c.dtype))  # JAX: int factors do not widen c
            for j in range(k - nu):
                cj = c[j] * fac[j] if nu > 0 else c[j]
                y = y * tt + cj
        y = y.reshape(x_shape + self.c.shape[2:])
        if not extrapolate:
            mask = np.logical_or(x > self.x[-1], x < self.x[0]).reshape(x_shape)
            y = np.where(mask.reshape(mask.shape + (1,) * (y.ndim - mask.ndim)), np.nan, y)
        if self.axis != 0:
            l = list(range(y.ndim))
            l = l[x_ndim:x_ndim + self.axis] + l[:x_ndim] + l[x_ndim + self.axis:]
            y = y.transpose(l)
        return y

    def derivative(self, nu=1):
        if nu < 0:
            return self.antiderivative(-nu)
        k = self.c.shape[0]
        if nu == 0:
            c2 = self.c.copy()
        else:
            fac = (polyder_coeffs(k, nu) if nu < k else np.zeros(0, dtype=np.int64)).astype(_real(self.c.dtype))
            c2 = self.c[:max(k - nu, 0)] * fac.reshape((-1,) + (1,) * (self.c.ndim - 1))
        if c2.shape[0] == 0:
            c2 = np.zeros((1,) + c2.shape[1:], dtype=c2.dtype)
        return self.construct_fast(c2, self.x, self.extrapolate, self.axis)

    def antiderivative(self, nu=1):
        if nu <= 0:
            return self.derivative(-nu)
        c2 = self.c.copy()
        dx = np.diff(self.x)
        for _ in range(nu):
            k = c2.shape[0]
            # jnp.polyint: divide by the new power, constant of integration 0
            div = np.arange(k, 0, -1).astype(_real(c2.dtype)).reshape((-1,) + (1,) * (c2.ndim - 1))
            c2 = np.concatenate([c2 / div, np.zeros((1,) + c2.shape[1:], dtype=c2.dtype)], axis=0)
            # value of each piece at its right end, accumulated into the constants
            z = np.zeros(c2.shape[1:], dtype=c2.dtype)
            dxr = dx.reshape((-1,) + (1,) * (c2.ndim - 2))
            for j in range(c2.shape[0]):
                z = z * dxr + c2[j]
            # the reference accumulates along `self.axis` of the ALREADY-moved array
            # (_ppoly.py:339), which is the breakpoint axis only for axis == 0
            c2[-1, 1:] += np.cumsum(z, axis=self.axis)[:-1]
        extrapolate = False if self.extrapolate == "periodic" else self.extrapolate
        return self.construct_fast(c2, self.x, extrapolate, self.axis)

    def integrate(self, a, b, extrapolate=None):
        if extrapolate is None:
            extrapolate = self.extrapolate
        integral = self.antiderivative(1)
        sign = 1 - 2 * (b < a)
        a, b = np.sort(np.array([a, b], dtype=self.x.dtype))
        tail = self.c.shape[2:]
        if extrapolate == "periodic":
            xs, xe = self.x[0], self.x[-1]
            period = xe - xs
            interval = b - a
            n_periods, left = np.divmod(interval, period)
            out = (integral(xe) - integral(xs)) * n_periods if n_periods > 0 else np.zeros(tail)
            a = xs + (a - xs) % period
            b = a + left
            if b <= xe:
                out = out + (integral(b) - integral(a))
            else:
                out = out + (integral(xe) - integral(a))
                out = out + (integral(xs + left + a - xe) - integral(xs))
        else:
            out = integral(b, extrapolate=extrapolate) - integral(a, extrapolate=extrapolate)
        return sign * np.asarray(out).reshape(tail)


def prepare_input(x, y, axis, dydx=None, check=True):
    """_ppoly.py:444-499."""
    x, y = map(asarray_inexact, (x, y))
    dx = np.diff(x)
    axis = axis % y.ndim
    if np.issubdtype(x.dtype, np.complexfloating):
        raise ValueError("`x` must contain real values.")
    if dydx is not None:
        dydx = asarray_inexact(dydx)
        if y.shape != dydx.shape:
            raise ValueError("The shapes of `y` and `dydx` must be identical.")
        dtype = np.promote_types(y.dtype, dydx.dtype)
        dydx = dydx.astype(dtype)
        y = y.astype(dtype)
    if check:
        if x.ndim != 1:
            raise ValueError("`x` must be 1-dimensional.")
        if x.shape[0] < 2:
            raise ValueError("`x` must contain at least 2 elements.")
        if x.shape[0] != y.shape[axis]:
            raise ValueError(f"The length of `y` along `axis`={axis} doesn't match the length of `x`")
        if not np.all(np.isfinite(x)):
            raise ValueError("`x` must contain only finite values.")
        if not np.all(np.isfinite(y)):
            raise ValueError("`y` must contain only finite values.")
        if dydx is not None and not np.all(np.isfinite(dydx)):
            raise ValueError("`dydx` must contain only finite values.")
        if np.any(dx <= 0):
            raise ValueError("`x` must be strictly increasing sequence.")
    return x, dx, y, axis, dydx


class CubicHermiteSpline(PPoly):
    """_ppoly.py:502-569: c = (A_CUBIC @ [y0, y1, d0 dx, d1 dx])[::-1] / dx**[3,2,1,0]."""

    def __init__(self, x, y, dydx, axis=0, extrapolate=None, check=True):
        if extrapolate is None:
            extrapolate = True
        x, dx, y, axis, dydx = prepare_input(x, y, axis, dydx, check)
        y = np.moveaxis(y, axis, 0)
        dydx = np.moveaxis(dydx, axis, 0)
        dxr = dx.reshape([dx.shape[0]] + [1] * (y.ndim - 1))
        F = np.stack([y[:-1], y[1:], dydx[:-1] * dxr, dydx[1:] * dxr], axis=0)     # (4, m, ...)
        dt = np.result_type(F.dtype, dx.dtype)
        coef = np.tensordot(A_CUBIC.astype(dt), F.astype(dt), axes=(1, 0))[::-1]    # t^3 first
        powers = np.arange(4)[::-1].reshape((4,) + (1,) * (y.ndim))
        c = coef / (dxr[None].astype(dt) ** powers)
        PPoly.__init__(self, np.moveaxis(np.moveaxis(c, 0, axis + 1), 0, axis + 1) if axis != 0 else c,
                       x, extrapolate=extrapolate, axis=axis, check=check)


class PchipInterpolator(CubicHermiteSpline):
    def __init__(self, x, y, axis=0, extrapolate=None, check=True):
        x, _, y, axis, _ = prepare_input(x, y, axis, check=check)
        dydx = approx_df(x, y, "monotonic", axis=axis)
        super().__init__(x, y, dydx, axis=axis, extrapolate=extrapolate, check=check)


class Akima1DInterpolator(CubicHermiteSpline):
    def __init__(self, x, y, axis=0, extrapolate=None, check=True):
        x, _, y, axis, _ = prepare_input(x, y, axis, check=check)
        t = approx_df(x, y, method="akima", axis=axis)
        super().__init__(x, y, t, axis=axis, extrapolate=extrapolate, check=check)


class CubicSpline(CubicHermiteSpline):
    def __init__(self, x, y, axis=0, bc_type="not-a-knot", extrapolate=None, check=True):
        x, _, y, axis, _ = prepare_input(x, y, axis, check=check)
        df = approx_df(x, y, "cubic2", axis, bc_type=bc_type)
        super().__init__(x, y, df, axis=axis, extrapolate=extrapolate, check=check)

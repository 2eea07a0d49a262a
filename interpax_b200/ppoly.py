"""Piecewise polynomials on the GPU: the host-side mirror of ``interpax/_ppoly.py``.

``PPoly``, ``CubicHermiteSpline``, ``PchipInterpolator``, ``Akima1DInterpolator`` and
``CubicSpline`` with the reference's signatures, argument checks and error texts
(_ppoly.py:50-441, 444-499, 502-809).  Every array operation that touches the coefficients or
the queries is a kernel behind ``libipx.so`` (``ipx_ppoly_eval_*``, ``ipx_ppoly_from_hermite_*``,
``ipx_ppoly_derivative_*``, ``ipx_ppoly_antiderivative_*``, and ``ipx_slopes_*`` for the knot
slopes); PyTorch only holds the device memory and moves axes.  There is no CPU fallback.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .api import (_NP2T, _T2NP, _device, _promote, _ptr, _slopes_dev, _stream, _suffix, _to_dev, errorif)

__all__ = ["PPoly", "CubicHermiteSpline", "PchipInterpolator", "Akima1DInterpolator", "CubicSpline"]

_MAX_ORDER = 64  # kPolyMaxOrder in csrc/ipx_ppoly.cu


def _real_of(dt: np.dtype) -> np.dtype:
    return np.dtype(np.float64) if np.dtype(dt) in (np.float64, np.complex128) else np.dtype(np.float32)


def _is_torch(*items) -> bool:
    return any(isinstance(a, torch.Tensor) for a in items)


class PPoly:
    """Piecewise polynomial in terms of coefficients and breakpoints; same contract as
    interpax.PPoly (_ppoly.py:50-441):  ``S = sum(c[m, i] * (xp - x[i])**(k-m) for m in range(k+1))``
    on ``x[i] <= xp < x[i+1]``.  ``c`` (k, m, ...), ``x`` (m+1,), ``extrapolate`` bool | "periodic",
    ``axis`` the interpolation axis."""

    def __init__(self, c, x, extrapolate=None, axis=0, check=True):
        want_torch = _is_torch(c, x)
        cdt, xdt = _promote([c]), _promote([x])
        cshape = tuple(c.shape) if hasattr(c, "shape") else np.shape(c)
        xshape = tuple(x.shape) if hasattr(x, "shape") else np.shape(x)
        errorif(len(cshape) < 2, ValueError, "Coefficients array must be at least 2-dimensional.")
        errorif(len(xshape) != 1, ValueError, "x must be 1-dimensional")
        errorif(xshape[0] < 2, ValueError, "at least 2 breakpoints are needed")
        axis = axis % (len(cshape) - 1)
        if extrapolate is None:
            extrapolate = True
        elif extrapolate != "periodic":
            extrapolate = bool(extrapolate)
        dev = _device()
        ct = _to_dev(c, cdt, dev)
        xt = _to_dev(x, xdt, dev)
        if axis != 0:
            # target shape (k, m, ...) from (..., k, m, ...): roll the two axes to the front
            ct = torch.movedim(ct, axis + 1, 0)
            ct = torch.movedim(ct, axis + 1, 0).contiguous()
        errorif(ct.shape[0] == 0, ValueError, "polynomial must be at least of order 0")
        errorif(ct.shape[1] != xt.shape[0] - 1, ValueError, "number of coefficients != len(x)-1")
        if check:
            errorif(bool((xt[1:] < xt[:-1]).any()), ValueError, "`x` must be strictly increasing sequence.")
        self._init_fast(ct, xt, extrapolate, axis, want_torch)

    def _init_fast(self, ct, xt, extrapolate, axis, want_torch):
        self._c = ct
        self._x = xt
        self._extrapolate = extrapolate
        self._axis = axis
        self._want_torch = want_torch
        self._ends = None

    # the reference exposes these as properties (_ppoly.py:146-164)
    @property
    def c(self):
        return self._c if self._want_torch else self._c.cpu().numpy()

    @property
    def x(self):
        return self._x if self._want_torch else self._x.cpu().numpy()

    @property
    def extrapolate(self):
        return self._extrapolate

    @property
    def axis(self):
        return self._axis

    @classmethod
    def construct_fast(cls, c, x, extrapolate=None, axis=0):
        """Construct without the checks (_ppoly.py:166-187); `c` must already be (k, m, ...)."""
        self = object.__new__(cls)
        want_torch = _is_torch(c, x)
        dev = _device()
        ct = _to_dev(c, _promote([c]), dev)
        xt = _to_dev(x, _promote([x]), dev)
        self._init_fast(ct, xt, True if extrapolate is None else extrapolate, axis, want_torch)
        return self

    @classmethod
    def _from_device(cls, ct, xt, extrapolate, axis, want_torch):
        self = object.__new__(cls)
        self._init_fast(ct, xt, extrapolate, axis, want_torch)
        return self

    # ------------------------------------------------------------------ kernels
    def _real_coeffs(self, rdt: np.dtype):
        """coefficients as a real (k, m, batch) tensor of dtype rdt (complex -> trailing pair)"""
        ct = self._c
        if ct.is_complex():
            ct = torch.view_as_real(ct)
        tdt = _NP2T[rdt]
        ct = ct.to(tdt).contiguous()
        k, m = ct.shape[0], ct.shape[1]
        batch = int(np.prod(ct.shape[2:], dtype=np.int64))
        return ct, k, m, batch

    def _piece_major(self, ct, k, m):
        """(m, k) copy of scalar-valued coefficients, built once per dtype (ipx_ppoly_pack_*)"""
        cache = self.__dict__.setdefault("_packed", {})
        if ct.dtype not in cache:
            out = torch.empty((m, k), dtype=ct.dtype, device=ct.device)
            fn = getattr(L.load(), f"ipx_ppoly_pack_{_suffix(ct)}")
            L.check(fn(_ptr(ct), k, m, _ptr(out), _stream()), "ipx_ppoly_pack")
            cache[ct.dtype] = out
        return cache[ct.dtype]

    def __call__(self, x, nu=0, extrapolate=None):
        """Evaluate the piecewise polynomial or its nu-th derivative (_ppoly.py:189-257).
        Result shape: the shape of ``x`` in place of the interpolation axis."""
        if extrapolate is None:
            extrapolate = self._extrapolate
        errorif(int(nu) != nu or nu < 0, ValueError, "Order of derivative must be a non-negative integer")
        want_torch = self._want_torch or _is_torch(x)
        cdt = _T2NP[self._c.dtype]
        xdt = np.result_type(_promote([x]), _T2NP[self._x.dtype])
        errorif(np.issubdtype(xdt, np.complexfloating), TypeError, "x must be real")
        odt = np.result_type(cdt, xdt)            # dtype of the reference's result
        rdt = _real_of(odt)                        # the kernels compute in its real part
        dev = self._c.device
        ct, k, m, batch = self._real_coeffs(rdt)
        errorif(k > _MAX_ORDER, NotImplementedError, f"polynomial order above {_MAX_ORDER}")
        xq = _to_dev(x, xdt, dev).to(_NP2T[rdt])
        x_shape, x_ndim = tuple(xq.shape), xq.ndim
        xq = xq.reshape(-1).contiguous()
        knots = self._x.to(_NP2T[rdt]).contiguous()
        mode = L.IPX_PPOLY_PERIODIC if extrapolate == "periodic" else (
            L.IPX_PPOLY_EXTRAPOLATE if extrapolate else L.IPX_PPOLY_NAN)
        out = torch.empty((xq.numel(), batch), dtype=ct.dtype, device=dev)
        layout = L.IPX_SOA
        if batch == 1 and k > 1 and xq.numel() >= 4 * m:
            # many queries on a scalar-valued polynomial: evaluate from the piece-major copy
            ct, layout = self._piece_major(ct, k, m), L.IPX_AOS
        fn = getattr(L.load(), f"ipx_ppoly_eval_{_suffix(ct)}")
        L.check(fn(_ptr(ct), layout, k, m, batch, _ptr(knots), _ptr(xq), xq.numel(), int(nu), mode, _ptr(out),
                   None, _stream()), "ipx_ppoly_eval")
        tail = tuple(self._c.shape[2:])
        if self._c.is_complex():
            out = torch.view_as_complex(out.reshape(x_shape + tail + (2,)))
        y = out.reshape(x_shape + tail)
        if y.dtype != _NP2T[np.dtype(odt)]:
            y = y.to(_NP2T[np.dtype(odt)])
        if self._axis != 0:
            # move the evaluated values to the interpolation axis (_ppoly.py:252-256)
            order = list(range(y.ndim))
            order = order[x_ndim:x_ndim + self._axis] + order[:x_ndim] + order[x_ndim + self._axis:]
            y = y.permute(order)
        return y if want_torch else y.cpu().numpy()

    def derivative(self, nu=1):
        """Piecewise polynomial of order k - nu representing the nu-th derivative
        (_ppoly.py:259-296); negative nu gives the antiderivative."""
        if nu < 0:
            return self.antiderivative(-nu)
        if nu == 0:
            return self._from_device(self._c.clone(), self._x, self._extrapolate, self._axis, self._want_torch)
        rdt = _real_of(_T2NP[self._c.dtype])
        ct, k, m, batch = self._real_coeffs(rdt)
        tail = tuple(ct.shape[2:])
        if nu >= k:
            c2 = torch.zeros((1, m) + tail, dtype=ct.dtype, device=ct.device)  # derivative of order 0 is zero
        else:
            c2 = torch.empty((k - nu, m) + tail, dtype=ct.dtype, device=ct.device)
            fn = getattr(L.load(), f"ipx_ppoly_derivative_{_suffix(ct)}")
            L.check(fn(_ptr(ct), k, m, batch, int(nu), _ptr(c2), _stream()), "ipx_ppoly_derivative")
        if self._c.is_complex():
            c2 = torch.view_as_complex(c2)
        return self._from_device(c2, self._x, self._extrapolate, self._axis, self._want_torch)

    def antiderivative(self, nu=1):
        """Piecewise polynomial of order k + nu representing the antiderivative, continuous
        across the breakpoints (_ppoly.py:298-346); periodic extrapolation becomes False."""
        if nu <= 0:
            return self.derivative(-nu)
        # the reference accumulates the constants along `self.axis` of the already-moved array
        # (_ppoly.py:339), which is the breakpoint axis only when axis == 0
        errorif(self._axis != 0, NotImplementedError, "antiderivative is implemented for axis=0")
        rdt = _real_of(_T2NP[self._c.dtype])
        ct, k, m, batch = self._real_coeffs(rdt)
        errorif(k + nu > _MAX_ORDER, NotImplementedError, f"polynomial order above {_MAX_ORDER}")
        tail = tuple(ct.shape[2:])
        knots = self._x.to(ct.dtype).contiguous()
        fn = getattr(L.load(), f"ipx_ppoly_antiderivative_{_suffix(ct)}")
        for _ in range(int(nu)):
            c2 = torch.empty((k + 1, m) + tail, dtype=ct.dtype, device=ct.device)
            L.check(fn(_ptr(ct), k, m, batch, _ptr(knots), _ptr(c2), _stream()), "ipx_ppoly_antiderivative")
            ct, k = c2, k + 1
        if self._c.is_complex():
            ct = torch.view_as_complex(ct)
        extrapolate = False if self._extrapolate == "periodic" else self._extrapolate
        return self._from_device(ct, self._x, extrapolate, self._axis, self._want_torch)

    def _end_points(self):
        if self._ends is None:
            e = self._x[[0, -1]].cpu().numpy()
            self._ends = (e[0], e[1])
        return self._ends

    def integrate(self, a, b, extrapolate=None):
        """Definite integral over [a, b] (_ppoly.py:348-419)."""
        if extrapolate is None:
            extrapolate = self._extrapolate
        integral = self.antiderivative(1)
        xdt = _T2NP[self._x.dtype]
        sign = 1 - 2 * int(b < a)
        a, b = np.sort(np.array([a, b], dtype=xdt))
        ev = lambda pts, **kw: PPoly.__call__(integral, torch.as_tensor(np.array(pts, dtype=xdt)), **kw)
        if extrapolate == "periodic":
            xs, xe = self._end_points()
            period = xe - xs
            n_periods, left = np.divmod(b - a, period)
            a = xs + (a - xs) % period
            b = a + left
            if b <= xe:
                v = ev([xe, xs, b, a])
                out = (v[0] - v[1]) * float(n_periods) if n_periods > 0 else torch.zeros_like(v[0])
                out = out + (v[2] - v[3])
            else:
                v = ev([xe, xs, a, xs + left + a - xe])
                out = (v[0] - v[1]) * float(n_periods) if n_periods > 0 else torch.zeros_like(v[0])
                out = out + (v[0] - v[2])
                out = out + (v[3] - v[1])
        else:
            v = ev([b, a], extrapolate=extrapolate)
            out = v[0] - v[1]
        out = sign * out.reshape(tuple(self._c.shape[2:]))
        return out if self._want_torch else out.cpu().numpy()

    def solve(self, y=0.0, discontinuity=True, extrapolate=None):
        raise NotImplementedError

    def roots(self, discontinuity=True, extrapolate=None):
        raise NotImplementedError

    def extend(self, c, x, right=True):
        raise NotImplementedError

    @classmethod
    def from_spline(cls, tck, extrapolate=None):
        raise NotImplementedError

    @classmethod
    def from_bernstein_basis(cls, bp, extrapolate=None):
        raise NotImplementedError


def _prepare_input(x, y, axis, dydx=None, check=True):
    """_ppoly.py:444-499 -> device tensors (x float64, y / dydx moved to axis 0 and contiguous),
    the normalised axis and whether the data are complex."""
    dev = _device()
    xdt = _promote([x])
    errorif(np.issubdtype(xdt, np.complexfloating), ValueError, "`x` must contain real values.")
    ydt = _promote([y])
    yshape = tuple(y.shape) if hasattr(y, "shape") else np.shape(y)
    axis = axis % len(yshape)
    if dydx is not None:
        dshape = tuple(dydx.shape) if hasattr(dydx, "shape") else np.shape(dydx)
        errorif(yshape != dshape, ValueError, "The shapes of `y` and `dydx` must be identical.")
        ydt = np.promote_types(ydt, _promote([dydx]))
    xt = _to_dev(x, np.dtype(np.float64), dev)  # x.astype(float), _ppoly.py:460
    # every later step mixes the data with the float64 knots (dydx * dx, approx_df(x, y)), so
    # the reference's coefficients are float64 / complex128 whatever y was; widen once here
    ydt = np.result_type(ydt, np.float64)
    yt = _to_dev(y, ydt, dev)
    dt = None if dydx is None else _to_dev(dydx, ydt, dev)
    if check:
        errorif(xt.ndim != 1, ValueError, "`x` must be 1-dimensional.")
        errorif(xt.shape[0] < 2, ValueError, "`x` must contain at least 2 elements.")
        errorif(xt.shape[0] != yt.shape[axis], ValueError,
                f"The length of `y` along `axis`={axis} doesn't match the length of `x`")
        errorif(not bool(torch.isfinite(xt).all()), ValueError, "`x` must contain only finite values.")
        fin = lambda t: bool(torch.isfinite(torch.view_as_real(t) if t.is_complex() else t).all())
        errorif(not fin(yt), ValueError, "`y` must contain only finite values.")
        errorif(dt is not None and not fin(dt), ValueError, "`dydx` must contain only finite values.")
        errorif(bool((xt[1:] <= xt[:-1]).any()), ValueError, "`x` must be strictly increasing sequence.")
    return xt, yt, axis, dt


class CubicHermiteSpline(PPoly):
    """Piecewise-cubic interpolator matching values and first derivatives; same contract as
    interpax.CubicHermiteSpline (_ppoly.py:502-569)."""

    def __init__(self, x, y, dydx, axis=0, extrapolate=None, check=True):
        want_torch = _is_torch(x, y, dydx)
        if extrapolate is None:
            extrapolate = True
        xt, yt, axis, dt = _prepare_input(x, y, axis, dydx, check)
        self._build(xt, yt, dt, axis, extrapolate, want_torch)

    def _build(self, xt, yt, dt, axis, extrapolate, want_torch):
        cplx = yt.is_complex()
        ym = torch.movedim(yt, axis, 0)
        dm = torch.movedim(dt, axis, 0)
        tail = tuple(ym.shape[1:])
        if cplx:
            ym, dm = torch.view_as_real(ym), torch.view_as_real(dm)
        # x is float64 (the reference casts it), so the coefficients are float64 / complex128
        ym = ym.to(torch.float64).contiguous()
        dm = dm.to(torch.float64).contiguous()
        n = xt.shape[0]
        batch = int(np.prod(ym.shape[1:], dtype=np.int64))
        c = torch.empty((4, n - 1) + tuple(ym.shape[1:]), dtype=torch.float64, device=ym.device)
        fn = L.load().ipx_ppoly_from_hermite_f64
        L.check(fn(_ptr(xt), n, _ptr(ym), _ptr(dm), batch, _ptr(c), _stream()), "ipx_ppoly_from_hermite")
        if cplx:
            c = torch.view_as_complex(c)
        assert tuple(c.shape[2:]) == tail
        self._init_fast(c, xt, extrapolate, axis, want_torch)


def _knot_slopes(xt, yt, method, axis, kwargs):
    """approx_df on device tensors; complex data as a trailing real pair."""
    if yt.is_complex():
        out = _slopes_dev(xt.to(_NP2T[_real_of(_T2NP[yt.dtype])]), torch.view_as_real(yt).contiguous(), method,
                          axis % yt.ndim, kwargs, True)
        return torch.view_as_complex(out)
    return _slopes_dev(xt.to(yt.dtype), yt.contiguous(), method, axis, kwargs)


class PchipInterpolator(CubicHermiteSpline):
    """PCHIP 1-D monotonic cubic interpolation; same contract as interpax.PchipInterpolator
    (_ppoly.py:572-652): slopes from approx_df(..., "monotonic")."""

    def __init__(self, x, y, axis=0, extrapolate=None, check=True):
        want_torch = _is_torch(x, y)
        xt, yt, axis, _ = _prepare_input(x, y, axis, check=check)
        dt = _knot_slopes(xt, yt, "monotonic", axis, {})
        self._build(xt, yt, dt, axis, True if extrapolate is None else extrapolate, want_torch)


class Akima1DInterpolator(CubicHermiteSpline):
    """Akima interpolator; same contract as interpax.Akima1DInterpolator (_ppoly.py:655-710)."""

    def __init__(self, x, y, axis=0, extrapolate=None, check=True):
        want_torch = _is_torch(x, y)
        xt, yt, axis, _ = _prepare_input(x, y, axis, check=check)
        dt = _knot_slopes(xt, yt, "akima", axis, {})
        self._build(xt, yt, dt, axis, True if extrapolate is None else extrapolate, want_torch)


class CubicSpline(CubicHermiteSpline):
    """Cubic spline data interpolator; same contract as interpax.CubicSpline (_ppoly.py:713-809):
    slopes from approx_df(..., "cubic2", bc_type=bc_type)."""

    def __init__(self, x, y, axis=0, bc_type="not-a-knot", extrapolate=None, check=True):
        want_torch = _is_torch(x, y)
        xt, yt, axis, _ = _prepare_input(x, y, axis, check=check)
        dt = _knot_slopes(xt, yt, "cubic2", axis, {"bc_type": bc_type})
        self._build(xt, yt, dt, axis, True if extrapolate is None else extrapolate, want_torch)

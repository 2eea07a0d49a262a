"""Derivatives with respect to the knots and to f, modelled on the
reference's TestAD (tests/test_interpolate.py:542-637): there jax.jacfwd and jax.jacrev of
interpNd / InterpolatorND with respect to the knot vector must agree with each other and with
finite differences.  Here: ipx.jvp_knots (forward) against ipx.vjp_knots (reverse) through
the dot-product identity, both against central finite differences of the ORACLE (and of the
kernels' own forward evaluation), on the reference's test setups and on harder ones
(non-uniform knots, periodic axes in both conventions, derivative orders, extrapolation)."""

import numpy as np
import pytest

import interpax_b200 as ipx
from oracle import interpax_np as O

pytestmark = pytest.mark.gpu

OFUN = (O.interp1d, O.interp2d, O.interp3d)
OCLS = (O.Interpolator1D, O.Interpolator2D, O.Interpolator3D)


def fd_knots(fun, knots, ax, h=1e-6):
    """Central finite differences d out / d knots[ax][k] -> (nq, n)."""
    cols = []
    for k in range(len(knots[ax])):
        kp = [v.copy() for v in knots]
        km = [v.copy() for v in knots]
        kp[ax][k] += h
        km[ax][k] -= h
        cols.append((fun(kp) - fun(km)) / (2 * h))
    return np.stack(cols, axis=-1)


def jacobians(qs, knots, f, method, functional, **kw):
    """(reverse, forward) Jacobians d out[q] / d knots[ax][k] from the kernels, per axis."""
    ndim = len(knots)
    nq = len(qs[0])
    rev = [np.zeros((nq, len(k))) for k in knots]
    for q in range(nq):
        cot = np.zeros(nq)
        cot[q] = 1.0
        g = ipx.vjp_knots(qs, knots, f, cot, method, functional=functional, **kw)
        for ax in range(ndim):
            rev[ax][q] = g["xyz"[ax]]
    fwd = [np.zeros((nq, len(k))) for k in knots]
    for ax in range(ndim):
        for k in range(len(knots[ax])):
            dots = [None] * ndim
            dots[ax] = np.zeros(len(knots[ax]))
            dots[ax][k] = 1.0
            fwd[ax][:, k] = ipx.jvp_knots(qs, knots, f, tuple(dots), None, method, functional=functional, **kw)
    return rev, fwd


@pytest.mark.parametrize("method", ["cubic", "cardinal", "catmull-rom"])
def test_reference_test_ad_interp1d(method):
    """tests/test_interpolate.py:543-569 for the methods with a local linear stencil."""
    xp = np.linspace(0, 2 * np.pi, 10)
    x = np.linspace(0, 2 * np.pi, 20)
    fp = np.sin(xp)
    kw = {"c": 0.3} if method == "cardinal" else {}
    for functional in (True, False):
        rev, fwd = jacobians((x,), (xp,), fp, method, functional, **kw)
        np.testing.assert_allclose(fwd[0], rev[0], rtol=1e-13, atol=1e-13)  # jacfwd == jacrev
        ofun = (lambda k: O.interp1d(x, k[0], fp, method, **kw)) if functional else (
            lambda k: O.Interpolator1D(k[0], fp, method, **kw)(x))
        jd = fd_knots(ofun, (xp,), 0)
        # (finite differences give nan at the end points, as the reference's test notes)
        np.testing.assert_allclose(fwd[0][1:-1], jd[1:-1], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("method", ["cubic", "cardinal"])
def test_reference_test_ad_interp2d_3d(method):
    """tests/test_interpolate.py:571-637: the Jacobian with respect to the first knot vector."""
    kw = {"c": 0.0} if method == "cardinal" else {}
    xp, yp = np.linspace(0, 4 * np.pi, 20), np.linspace(0, 2 * np.pi, 20)
    x = y = np.linspace(0, 2 * np.pi, 30)
    xxp, yyp = np.meshgrid(xp, yp, indexing="ij")
    fp = np.sin(xxp) * np.cos(yyp)
    rev, fwd = jacobians((x, y), (xp, yp), fp, method, True, **kw)
    np.testing.assert_allclose(fwd[0], rev[0], rtol=1e-13, atol=1e-13)
    jd = fd_knots(lambda k: O.interp2d(x, y, k[0], k[1], fp, method, **kw), (xp, yp), 0)
    np.testing.assert_allclose(fwd[0][1:-1], jd[1:-1], rtol=1e-6, atol=1e-6)
    xp, yp, zp = np.linspace(0, np.pi, 10), np.linspace(0, 2 * np.pi, 15), np.linspace(0, 1, 12)
    x, y, z = np.linspace(0, np.pi, 13), np.linspace(0, 2 * np.pi, 13), np.linspace(0, 1, 13)
    xxp, yyp, zzp = np.meshgrid(xp, yp, zp, indexing="ij")
    fp = np.sin(xxp) * np.cos(yyp) * zzp**2
    for functional in (True, False):
        rev, fwd = jacobians((x, y, z), (xp, yp, zp), fp, method, functional, **kw)
        for ax in range(3):
            np.testing.assert_allclose(fwd[ax], rev[ax], rtol=1e-12, atol=1e-12)
        ofun = (lambda k: O.interp3d(x, y, z, *k, fp, method, **kw)) if functional else (
            lambda k: O.Interpolator3D(*k, fp, method, **kw)(x, y, z))
        for ax in range(3):
            jd = fd_knots(ofun, (xp, yp, zp), ax)
            np.testing.assert_allclose(fwd[ax][1:-1], jd[1:-1], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_knot_and_table_derivatives_general(ndim):
    """Non-uniform knots, periodic axes in both of the reference's conventions, derivative
    orders, extrapolation with fills: the reverse rule is the transpose of the forward rule
    (random cotangent / tangents), and both match central differences of the oracle with
    respect to every knot vector and to f."""
    rng = np.random.default_rng(300 + ndim)
    ns = [9, 7, 8][:ndim]
    nq = 300
    for trial in range(6):
        functional = trial % 2 == 0
        method = ["cubic", "cardinal", "catmull-rom"][trial % 3]
        kw = {"c": 0.25} if method == "cardinal" else {}
        knots = [np.sort(rng.uniform(0.05, 2.0, n)) for n in ns]
        for k in knots:
            k[0] = 0.05
        f = rng.normal(size=ns)
        per_axes = [bool((trial >> 1) & 1) and ax != ndim - 1 for ax in range(ndim)] if ndim > 1 else [trial >= 4]
        period = tuple(2.3 if p else None for p in per_axes)
        period = period[0] if ndim == 1 else period
        deriv = tuple(int(v) for v in rng.integers(0, 3, ndim)) if trial % 3 == 1 else (0,) * ndim
        dfun = deriv[0] if ndim == 1 else deriv
        extrap = [True, (True, 0.5), False][trial % 3]
        # queries away from the knots (the map is only piecewise smooth in the knots there)
        qs = [rng.uniform(-0.3, 2.4, nq) for _ in range(ndim)]
        what = f"nd={ndim} {method} functional={functional} per={period} d={deriv} ex={extrap}"
        if functional:
            ofun = lambda k, ff=f: OFUN[ndim - 1](*qs, *k, ff, method, dfun, extrap, period, **kw)
        else:
            ofun = lambda k, ff=f: OCLS[ndim - 1](*k, ff, method, extrap, period, **kw)(*qs, *deriv)
        base = ofun(knots)
        ok = np.isfinite(base)
        cot = rng.normal(size=nq)
        g = ipx.vjp_knots(qs, knots, f, cot, method, dfun, extrap, period, functional=functional, **kw)
        kd = [rng.normal(size=n) for n in ns]
        fd = rng.normal(size=ns)
        tan = ipx.jvp_knots(qs, knots, f, tuple(kd), fd, method, dfun, extrap, period, functional=functional, **kw)
        # transpose identity: <cot, J v> == <J^T cot, v>
        lhs = float(np.dot(cot[ok], tan[ok]))
        rhs = sum(float(np.dot(g["xyz"[ax]], kd[ax])) for ax in range(ndim)) + float(np.sum(g["f"] * fd))
        assert np.all(tan[~ok] == 0)
        assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs)), what
        # directional finite difference of the oracle along (kd, fd)
        h = 1e-6
        kp = [k + h * d for k, d in zip(knots, kd)]
        km = [k - h * d for k, d in zip(knots, kd)]
        num = (ofun(kp, f + h * fd) - ofun(km, f - h * fd)) / (2 * h)
        smooth = ok & np.isfinite(num)
        # drop queries whose cell changes under the perturbation (a knot crosses the query)
        for ax in range(ndim):
            P = period if ndim == 1 else period[ax]
            for kk in (kp, km, knots):
                kv, qv = (np.sort(np.mod(kk[ax], P)), np.mod(qs[ax], P)) if P is not None else (kk[ax], qs[ax])
                cell = np.searchsorted(kv, qv, side="right")
                if kk is kp:
                    ref_cell = cell
                smooth &= cell == ref_cell
        scale = np.abs(num[smooth]).max()
        np.testing.assert_allclose(tan[smooth], num[smooth], rtol=2e-5, atol=2e-6 * scale, err_msg=what)


BC_PAIRS = ["not-a-knot", "natural", "clamped", ("not-a-knot", "natural"), ("clamped", "not-a-knot"),
            ((1, 0.7), (2, -0.4)), ((2, 0.3), "not-a-knot")]


@pytest.mark.parametrize("method", ["cubic", "cardinal", "catmull-rom", "cubic2"])
def test_approx_df_vjp_is_the_transpose_of_approx_df(method):
    """approx_df is affine in f for these methods (_fd_derivs.py:79-101, 160-300, 303-324); its
    reverse rule must satisfy <g, D v> == <D^T g, v> with D v taken from the ORACLE, for every axis
    position, non-uniform knots, every boundary-condition pair of cubic2, a duplicated knot, and
    must reproduce the dense transpose column by column on a small case."""
    rng = np.random.default_rng(77)
    bcs = BC_PAIRS if method == "cubic2" else [None]
    for bc in bcs:
        for shape, axis in (((11,), 0), ((4, 9, 3), 1), ((5, 8), -1), ((7, 6), 0), ((3,), 0) if bc != "not-a-knot" else ((4,), 0)):
            n = shape[axis]
            x = np.sort(rng.uniform(0.0, 3.0, n))
            if n >= 9:
                x[5] = x[4]  # duplicate knot: the row is neutralised (_fd_derivs.py:281-286)
            kw = {"c": 0.35} if method == "cardinal" else {}
            kw0 = dict(kw)
            if bc is not None:
                def full(b, zero):
                    if isinstance(b, tuple) and not isinstance(b[0], (str, tuple)):
                        vshape = shape[:axis % len(shape)] + shape[axis % len(shape) + 1:]
                        return (b[0], np.zeros(vshape) if zero else np.full(vshape, b[1]))
                    return b
                pair = (bc, bc) if isinstance(bc, str) else bc
                kw["bc_type"] = tuple(full(b, False) for b in pair)
                kw0["bc_type"] = tuple(full(b, True) for b in pair)
            v = rng.normal(size=shape)
            g = rng.normal(size=shape)
            with np.errstate(all="ignore"):
                Dv = O.approx_df(x, v, method, axis, **kw0)  # zero boundary values: the linear part
            got = ipx.approx_df_vjp(x, g, method, axis, **kw)
            assert got.shape == shape
            lhs, rhs = float(np.sum(g * Dv)), float(np.sum(got * v))
            assert abs(lhs - rhs) <= 1e-11 * max(1.0, abs(lhs), np.abs(g).max() * np.abs(Dv).max() * v.size), (method, bc, shape, axis)
    # dense transpose, column by column
    n = 8
    x = np.sort(rng.uniform(0, 2, n))
    kw = {"c": 0.2} if method == "cardinal" else {}
    D = np.stack([O.approx_df(x, e, method, 0, **kw) for e in np.eye(n)], axis=1)  # D[k, j]
    for k in range(n):
        np.testing.assert_allclose(ipx.approx_df_vjp(x, np.eye(n)[k], method, 0, **kw), D[k], rtol=1e-12, atol=1e-13)
    # float32 runs too
    got32 = ipx.approx_df_vjp(x.astype(np.float32), np.eye(n, dtype=np.float32)[3], method, 0, **kw)
    assert got32.dtype == np.float32
    np.testing.assert_allclose(got32, D[3], rtol=2e-4, atol=2e-4 * np.abs(D).max())


def test_approx_df_vjp_rejects_what_is_not_linear():
    x = np.linspace(0, 1, 6)
    for m in ("monotonic", "akima"):
        with pytest.raises(NotImplementedError):
            ipx.approx_df_vjp(x, np.ones(6), m)


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_vjp_f_is_the_transpose_of_the_evaluation(ndim):
    """Gradient with respect to f with the slope chain inside, for every method linear in f
    (cubic2 included: transposed tridiagonal solves along each axis of the chain), both periodic
    conventions, derivative orders, extrapolation fills: <cot, J v> == <J^T cot, v> with J v from
    the ORACLE's own evaluation of the tangent table (the map is linear in f), and for the local
    stencils the same gradient as ipx.vjp_knots gives."""
    rng = np.random.default_rng(500 + ndim)
    ns = [9, 7, 8][:ndim]
    nq = 400
    for trial in range(12):
        functional = trial % 2 == 0
        method = ["cubic2", "cubic", "linear", "cardinal", "cubic2", "catmull-rom"][trial % 6]
        kw = {"c": 0.25} if method == "cardinal" else {}
        if method == "cubic2":
            kw["bc_type"] = ["not-a-knot", "natural", ("clamped", "not-a-knot")][(trial // 2) % 3]
        knots = [np.sort(rng.uniform(0.05, 2.0, n)) for n in ns]
        for k in knots:
            k[0] = 0.05
        per_axes = [bool((trial >> 1) & 1) and ax != ndim - 1 for ax in range(ndim)] if ndim > 1 else [trial % 4 >= 2]
        period = tuple(2.3 if p else None for p in per_axes)
        period = period[0] if ndim == 1 else period
        deriv = tuple(int(v) for v in rng.integers(0, 3, ndim)) if trial % 3 == 1 else (0,) * ndim
        dfun = deriv[0] if ndim == 1 else deriv
        extrap = [True, (True, 0.5), False][trial % 3]
        qs = [rng.uniform(-0.3, 2.4, nq) for _ in range(ndim)]
        what = f"nd={ndim} {method} functional={functional} per={period} d={deriv} ex={extrap} {kw}"
        v = rng.normal(size=ns)
        with np.errstate(all="ignore"):
            if functional:
                Jv = OFUN[ndim - 1](*qs, *knots, v, method, dfun, extrap, period, **kw)
                J0 = OFUN[ndim - 1](*qs, *knots, np.zeros(ns), method, dfun, extrap, period, **kw)
            else:
                Jv = OCLS[ndim - 1](*knots, v, method, extrap, period, **kw)(*qs, *deriv)
                J0 = OCLS[ndim - 1](*knots, np.zeros(ns), method, extrap, period, **kw)(*qs, *deriv)
        Jv = Jv - J0  # a numeric fill value is a constant, not part of the linear map
        ok = np.isfinite(Jv)
        cot = rng.normal(size=nq)
        cot[~ok] = 0.0
        g = ipx.vjp_f(qs, knots, ns, cot, method, dfun, extrap, period, functional=functional, **kw)
        assert g.shape == tuple(ns)
        lhs, rhs = float(np.dot(cot[ok], Jv[ok])), float(np.sum(g * v))
        assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs), np.abs(Jv[ok]).max() * np.abs(cot).max()), what
        if method in ("cubic", "cardinal", "catmull-rom") and not (
                not functional and any(per_axes)):
            g2 = ipx.vjp_knots(qs, knots, rng.normal(size=ns), cot, method, dfun, extrap, period,
                               functional=functional, **kw)["f"]
            np.testing.assert_allclose(g, g2, rtol=1e-10, atol=1e-11 * max(1.0, np.abs(g2).max()), err_msg=what)


def test_vjp_f_batched_tables():
    """Trailing batch dims of f: the gradient has the shape of f and each batch element its own."""
    rng = np.random.default_rng(9)
    x, y = np.sort(rng.uniform(0, 1, 10)), np.sort(rng.uniform(0, 1, 12))
    qs = [rng.uniform(0, 1, 50), rng.uniform(0, 1, 50)]
    fshape = (10, 12, 3)
    cot = rng.normal(size=(50, 3))
    g = ipx.vjp_f(qs, (x, y), fshape, cot, "cubic2")
    for b in range(3):
        gb = ipx.vjp_f(qs, (x, y), (10, 12), cot[:, b], "cubic2")
        np.testing.assert_allclose(g[..., b], gb, rtol=1e-12, atol=1e-13)


def rev_jacobian(qs, knots, f, method, functional, **kw):
    """d out[q] / d knots[ax][k] by reverse mode, one unit cotangent per query."""
    nq = len(qs[0])
    rev = [np.zeros((nq, len(k))) for k in knots]
    for q in range(nq):
        cot = np.zeros(nq)
        cot[q] = 1.0
        g = ipx.vjp_knots(qs, knots, f, cot, method, functional=functional, **kw)
        for ax in range(len(knots)):
            rev[ax][q] = g["xyz"[ax]]
    return rev


@pytest.mark.parametrize("method", ["cubic2", "monotonic"])
def test_reference_test_ad_interp1d_chained_methods(method):
    """tests/test_interpolate.py:543-569 for the two methods of its list that are evaluated from
    tables: the reverse-mode Jacobian with respect to the knots against central differences of the
    oracle, functional and class form (which coincide without a period, as the reference asserts)."""
    xp = np.linspace(0, 2 * np.pi, 10)
    x = np.linspace(0, 2 * np.pi, 20)
    fp = np.sin(xp)
    revs = []
    for functional in (True, False):
        rev = rev_jacobian((x,), (xp,), fp, method, functional)[0]
        ofun = (lambda k: O.interp1d(x, k[0], fp, method)) if functional else (
            lambda k: O.Interpolator1D(k[0], fp, method)(x))
        jd = fd_knots(ofun, (xp,), 0)
        np.testing.assert_allclose(rev[1:-1], jd[1:-1], rtol=1e-6, atol=1e-6)
        revs.append(rev)
    np.testing.assert_allclose(revs[0], revs[1], rtol=1e-14, atol=1e-14)


def test_reference_test_ad_interp2d_3d_cubic2():
    """tests/test_interpolate.py:571-637 for method cubic2: Jacobian with respect to the knot vectors."""
    xp, yp = np.linspace(0, 4 * np.pi, 20), np.linspace(0, 2 * np.pi, 20)
    x = y = np.linspace(0, 2 * np.pi, 30)
    xxp, yyp = np.meshgrid(xp, yp, indexing="ij")
    fp = np.sin(xxp) * np.cos(yyp)
    rev = rev_jacobian((x, y), (xp, yp), fp, "cubic2", True)
    jd = fd_knots(lambda k: O.interp2d(x, y, k[0], k[1], fp, "cubic2"), (xp, yp), 0)
    np.testing.assert_allclose(rev[0][1:-1], jd[1:-1], rtol=1e-6, atol=1e-6)
    xp, yp, zp = np.linspace(0, np.pi, 10), np.linspace(0, 2 * np.pi, 15), np.linspace(0, 1, 12)
    x, y, z = np.linspace(0, np.pi, 13), np.linspace(0, 2 * np.pi, 13), np.linspace(0, 1, 13)
    xxp, yyp, zzp = np.meshgrid(xp, yp, zp, indexing="ij")
    fp = np.sin(xxp) * np.cos(yyp) * zzp**2
    for functional in (True, False):
        rev = rev_jacobian((x, y, z), (xp, yp, zp), fp, "cubic2", functional)
        ofun = (lambda k: O.interp3d(x, y, z, *k, fp, "cubic2")) if functional else (
            lambda k: O.Interpolator3D(*k, fp, "cubic2")(x, y, z))
        for ax in range(3):
            jd = fd_knots(ofun, (xp, yp, zp), ax)
            np.testing.assert_allclose(rev[ax][1:-1], jd[1:-1], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_chained_knot_and_table_gradients_general(ndim):
    """cubic2 (every boundary-condition kind, boundary values included) and monotonic on non-uniform
    knots, periodic axes in both conventions, derivative orders, fills: the reverse rule against a
    directional central difference of the oracle along random knot and table perturbations."""
    rng = np.random.default_rng(900 + ndim)
    ns = [9, 7, 8][:ndim]
    nq = 300
    for trial in range(10):
        functional = trial % 2 == 0
        method = ["cubic2", "monotonic", "cubic2", "monotonic-0", "cubic2"][trial % 5]
        kw = {}
        knots = [np.sort(rng.uniform(0.05, 2.0, n)) for n in ns]
        for k in knots:
            k[0] = 0.05
        # (smooth data: monotonic switches branches where secants change sign)
        f = rng.normal(size=ns) if method == "cubic2" else np.cumsum(rng.uniform(0.2, 1.0, size=ns), axis=0)
        if method == "cubic2":
            bcs = ["not-a-knot", "natural", ("clamped", "not-a-knot"), None, "not-a-knot"]
            bc = bcs[(trial // 2) % 5]
            if bc is None:
                vshape0 = tuple(ns[1:])  # the chain's first link runs along axis 0 only in 1-D
                bc = ((1, rng.normal(size=vshape0)), (2, rng.normal(size=vshape0))) if ndim == 1 else "natural"
            kw["bc_type"] = bc
        per_axes = [bool((trial >> 1) & 1) and ax != ndim - 1 for ax in range(ndim)] if ndim > 1 else [trial % 4 >= 2]
        if method != "cubic2" and any(per_axes):
            per_axes = [False] * ndim  # monotone data is not periodic
        period = tuple(2.3 if p else None for p in per_axes)
        period = period[0] if ndim == 1 else period
        deriv = tuple(int(v) for v in rng.integers(0, 3, ndim)) if trial % 3 == 1 else (0,) * ndim
        dfun = deriv[0] if ndim == 1 else deriv
        extrap = [True, (True, 0.5), False][trial % 3]
        qs = [rng.uniform(-0.3, 2.4, nq) for _ in range(ndim)]
        what = f"nd={ndim} {method} functional={functional} per={period} d={deriv} ex={extrap} {list(kw)}"
        if functional:
            ofun = lambda k, ff: OFUN[ndim - 1](*qs, *k, ff, method, dfun, extrap, period, **kw)
        else:
            ofun = lambda k, ff: OCLS[ndim - 1](*k, ff, method, extrap, period, **kw)(*qs, *deriv)
        with np.errstate(all="ignore"):
            base = ofun(knots, f)
        ok = np.isfinite(base)
        kd = [rng.normal(size=n) for n in ns]
        fd = rng.normal(size=ns)
        h = 1e-6
        kp = [k + h * d for k, d in zip(knots, kd)]
        km = [k - h * d for k, d in zip(knots, kd)]
        with np.errstate(all="ignore"):
            num = (ofun(kp, f + h * fd) - ofun(km, f - h * fd)) / (2 * h)
        smooth = ok & np.isfinite(num)
        for ax in range(ndim):
            P = period if ndim == 1 else period[ax]
            for kk in (kp, km, knots):
                kv, qv = (np.sort(np.mod(kk[ax], P)), np.mod(qs[ax], P)) if P is not None else (kk[ax], qs[ax])
                cell = np.searchsorted(kv, qv, side="right")
                if kk is kp:
                    ref_cell = cell
                smooth &= cell == ref_cell
        cot = rng.normal(size=nq)
        cot[~smooth] = 0.0
        g = ipx.vjp_knots(qs, knots, f, cot, method, dfun, extrap, period, functional=functional, **kw)
        lhs = float(np.dot(cot[smooth], num[smooth]))
        rhs = sum(float(np.dot(g["xyz"[ax]], kd[ax])) for ax in range(ndim)) + float(np.sum(g["f"] * fd))
        scale = max(1.0, float(np.abs(cot[smooth] * num[smooth]).sum()))
        assert abs(lhs - rhs) <= 2e-5 * scale, (what, lhs, rhs)


@pytest.mark.parametrize("method", ["cubic2", "monotonic"])
def test_chained_forward_mode_equals_reverse_mode(method):
    """jacfwd == jacrev for the chained methods (tests/test_interpolate.py:564-566, 632-634), 1-D on
    the reference's setup and 3-D with a periodic axis, plus f tangents through the transpose identity."""
    xp = np.linspace(0, 2 * np.pi, 10)
    x = np.linspace(0, 2 * np.pi, 20)
    fp = np.sin(xp)
    for functional in (True, False):
        rev = rev_jacobian((x,), (xp,), fp, method, functional)[0]
        fwd = np.zeros_like(rev)
        for k in range(len(xp)):
            dots = np.zeros(len(xp))
            dots[k] = 1.0
            fwd[:, k] = ipx.jvp_knots((x,), (xp,), fp, (dots,), None, method, functional=functional)
        np.testing.assert_allclose(fwd, rev, rtol=1e-13, atol=1e-13)
    rng = np.random.default_rng(4242)
    ns = [8, 7, 9]
    knots = [np.sort(rng.uniform(0.05, 2.0, n)) for n in ns]
    f = rng.normal(size=ns) if method == "cubic2" else np.cumsum(rng.uniform(0.2, 1.0, size=ns), axis=0)
    qs = [rng.uniform(-0.3, 2.4, 500) for _ in range(3)]
    for functional in (True, False):
        for period, extrap, deriv in (((None, None, None), (True, 0.5), (0, 0, 0)),
                                      ((2.3, None, None), True, (1, 0, 2)) if method == "cubic2" else
                                      ((None, None, None), False, (0, 1, 0))):
            kw = {"bc_type": ("natural", "not-a-knot")} if method == "cubic2" else {}
            cot = rng.normal(size=500)
            kd = [rng.normal(size=n) for n in ns]
            fd = rng.normal(size=ns)
            g = ipx.vjp_knots(qs, knots, f, cot, method, deriv, extrap, period, functional=functional, **kw)
            tan = ipx.jvp_knots(qs, knots, f, tuple(kd), fd, method, deriv, extrap, period, functional=functional, **kw)
            assert np.all(np.isfinite(tan))
            lhs = float(np.dot(cot, tan))
            rhs = sum(float(np.dot(g["xyz"[ax]], kd[ax])) for ax in range(3)) + float(np.sum(g["f"] * fd))
            assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs)), (method, functional, period, lhs, rhs)

"""GPU parity tests: the CUDA path (through the public API -> C ABI -> libipx.so) against the
CPU oracle (oracle/interpax_np.py) on the same seeded inputs.

Bars (BASELINE.json north_star):
  * interval indices: bit-exact against searchsorted(side="right") + clip;
  * values and derivatives: rtol 1e-12 in float64, 1e-5 in float32, each paired with
    atol = rtol * max|oracle result| (the reference's tests always pair rtol with an atol;
    a pure rtol is meaningless at zero crossings of the interpolant).
"""

import ctypes as C
import itertools

import numpy as np
import pytest
import torch

import interpax_b200 as ipx
from interpax_b200 import _lib as L
from interpax_b200 import api
from oracle import interpax_np as O

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5,
        np.dtype(np.complex128): 1e-12, np.dtype(np.complex64): 1e-5}
ALL_METHODS = O.METHODS


# reported by conftest at the end of the run
HATCH = {"elements_checked": 0, "accepted_via_f64_truth": 0, "accepted_via_oracle_noise": 0}


def assert_parity(got, want, what="", truth=None):
    """|got - want| <= rtol*|want| + rtol*max|want| element-wise.

    `truth` (optional, float32 cases only) is the same function evaluated by the oracle in
    float64 on the same float32 inputs.  Far outside the knots the cubic is ill-conditioned
    (weights grow like |t|^3 and cancel) and the float32 ORACLE is itself off by more than
    1e-5; its dense 64x64 form loses more digits than the separable form the kernels use.
    An element that misses the plain tolerance against the float32 oracle is therefore still
    accepted if (a) the KERNEL is within the same tolerance of the float64 evaluation -- the bound
    the kernel can be held to whatever the oracle's own float32 noise is -- or, where float32
    itself cannot do that (second derivatives of a float32 cubic2 solve far outside the knots),
    (b) the kernel is no further from the float64 evaluation than the tolerance plus twice the
    float32 oracle's own worst distance from it on the same query set.  Both kinds are counted
    (HATCH) and the counts are printed at the end of the test run.
    """
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    rtol = RTOL[np.dtype(want.dtype)]
    fin = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want)), what
    assert np.array_equal(np.isinf(got), np.isinf(want)), what
    if not fin.any():
        return
    scale = np.abs(want[fin]).max()
    tol = rtol * np.abs(want[fin]) + rtol * scale
    bad = np.abs(got[fin] - want[fin]) > tol
    HATCH["elements_checked"] += int(fin.sum())
    if bad.any() and truth is not None:
        truth = np.asarray(truth)[fin]
        tscale = np.abs(truth).max()
        dist = np.abs(got[fin].astype(truth.dtype) - truth)
        ttol = rtol * np.abs(truth) + rtol * tscale
        ok_vs_truth = dist <= ttol
        HATCH["accepted_via_f64_truth"] += int((bad & ok_vs_truth).sum())
        bad &= ~ok_vs_truth
        oracle_noise = np.abs(want[fin].astype(truth.dtype) - truth).max()
        ok_noise = dist <= ttol + 2 * oracle_noise
        HATCH["accepted_via_oracle_noise"] += int((bad & ok_noise).sum())
        bad &= ~ok_noise
    if bad.any():
        np.testing.assert_allclose(got[fin], want[fin], rtol=rtol, atol=rtol * scale, err_msg=what)


def upcast(a):
    if isinstance(a, np.ndarray) and a.dtype == np.float32:
        return a.astype(np.float64)
    return a


def knots_nonuniform(rng, n, lo, hi, dtype):
    x = np.sort(rng.uniform(lo, hi, n))
    x[0], x[-1] = lo, hi
    return x.astype(dtype)


# ------------------------------------------------------------------ C1: the README example
def test_readme_example_c1():
    xp = np.linspace(0, 2 * np.pi, 100)
    xq = np.linspace(0, 2 * np.pi, 10000)
    fp = np.sin(xp)
    fq, idx = api._interp(1, (xq,), (xp,), fp, "cubic", 0, False, None, {}, return_index=True)
    np.testing.assert_allclose(fq, np.sin(xq), rtol=1e-6, atol=1e-5)  # README.rst:32-44
    assert_parity(fq, O.interp1d(xq, xp, fp, method="cubic"))
    assert np.array_equal(idx[0], O.bin_index(xp, xq).astype(np.int32))


# ------------------------------------------------------------------ interval indices, bit-exact
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bin_index_bit_exact(dtype):
    rng = np.random.default_rng(1)
    cases = []
    cases.append(np.linspace(0, 1, 1000).astype(dtype))                      # uniform
    cases.append(knots_nonuniform(rng, 777, -3, 5, dtype))                   # non-uniform
    x = knots_nonuniform(rng, 300, 0, 1, dtype)
    x[50:60] = x[50]                                                         # duplicated knots
    x[-3:] = x[-1]
    cases.append(x)
    cases.append(np.array([0.0, 1.0], dtype=dtype))                          # two knots
    cases.append(np.concatenate([np.linspace(0, 1e-3, 500), np.linspace(1, 100, 500)]).astype(dtype))
    cases.append(np.zeros(16, dtype=dtype))                                  # all equal
    for x in cases:
        span = float(x[-1] - x[0]) or 1.0
        q = rng.uniform(x[0] - 0.1 * span, x[-1] + 0.1 * span, 20000).astype(dtype)
        special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, x[0], x[-1], np.nextafter(x[0], -np.inf),
                            np.nextafter(x[-1], np.inf)], dtype=dtype)
        q = np.concatenate([q, special, x, np.nextafter(x, dtype(np.inf)), np.nextafter(x, dtype(-np.inf))])
        got = ipx.bin_index(x, q)
        assert got.dtype == np.int32
        assert np.array_equal(got, O.bin_index(x, q)), len(x)


def test_bin_index_million_knots():
    rng = np.random.default_rng(2)
    x = np.linspace(0, 1, 1_000_000)
    xj = np.sort(x + 0.3 * (x[1] - x[0]) * rng.uniform(-1, 1, x.size))      # C2's jittered variant
    q = rng.uniform(-0.01, 1.01, 1 << 20)
    for knots in (x, xj):
        assert np.array_equal(ipx.bin_index(knots, q), O.bin_index(knots, q))


# ------------------------------------------------------------------ slope stencils
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", list(O.CUBIC_METHODS) + ["linear"])
def test_slopes_vs_oracle(method, dtype):
    rng = np.random.default_rng(3)
    for shape, axis in [((40,), 0), ((33, 5), 0), ((4, 29), 1), ((3, 17, 6), 1), ((5, 4, 21), 2),
                        ((2,), 0), ((3,), 0), ((4,), 0), ((2, 3), 0), ((3, 2), 0)]:
        n = shape[axis]
        x = knots_nonuniform(rng, n, 0, 2, dtype)
        f = rng.normal(size=shape).astype(dtype)
        kw = {"c": 0.3} if method == "cardinal" else {}
        want = O.approx_df(x, f, method, axis, **kw)
        got = ipx.approx_df(x, f, method, axis, **kw)
        assert_parity(got, want.astype(dtype), f"{method} {shape} axis={axis}")


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_slopes_complex_data(dtype):
    """Complex `f` (the reference's tests run every method on complex data,
    tests/test_interpolate.py:28-81): all methods but akima act on the two parts separately;
    akima weighs with moduli of complex differences (_fd_derivs.py:414)."""
    rng = np.random.default_rng(5)
    rdt = np.float32 if dtype == np.complex64 else np.float64
    for shape, axis in [((40,), 0), ((33, 5), 0), ((4, 29), 1), ((3, 17, 6), 1), ((5, 4, 21), -1),
                        ((2,), 0), ((3,), 0), ((4,), 0), ((3, 2), 0)]:
        n = shape[axis]
        x = knots_nonuniform(rng, n, 0, 2, rdt)
        f = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(dtype)
        if n >= 12:
            sl = [slice(None)] * len(shape)
            sl[axis] = slice(5, 9)
            f[tuple(sl)] = 0.25 - 1.5j  # flat stretch: degenerate akima weights -> 0.5 (m1 + m4)
        for method in O.CUBIC_METHODS:
            want = O.approx_df(x, f, method, axis)
            got = ipx.approx_df(x, f, method, axis)
            assert got.dtype == dtype
            assert_parity(got, want.astype(dtype), f"{method} {shape} axis={axis}")
    # 2-D: the slope chain fx, fy, fxy on complex data, functional and class form
    xk, yk = knots_nonuniform(rng, 11, 0, 1, rdt), knots_nonuniform(rng, 9, 0, 2, rdt)
    f = (rng.normal(size=(11, 9)) + 1j * rng.normal(size=(11, 9))).astype(dtype)
    xq, yq = rng.uniform(0, 1, 300).astype(rdt), rng.uniform(0, 2, 300).astype(rdt)
    want = O.interp2d(xq, yq, xk, yk, f, "akima")
    assert_parity(ipx.interp2d(xq, yq, xk, yk, f, "akima"), want, "interp2d akima complex")
    assert_parity(ipx.Interpolator2D(xk, yk, f, "akima")(xq, yq), O.Interpolator2D(xk, yk, f, "akima")(xq, yq),
                  "Interpolator2D akima complex")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_slopes_duplicate_knots_and_flat_data(dtype):
    rng = np.random.default_rng(4)
    x = knots_nonuniform(rng, 30, 0, 1, dtype)
    x[10] = x[11]
    x[20:23] = x[20]
    f = rng.normal(size=(30, 3)).astype(dtype)
    f[5:9] = f[5]  # flat stretch: monotonic zero slopes, akima degenerate weights
    for method in O.CUBIC_METHODS:
        if method == "cubic2":
            continue  # a duplicated knot makes the reference's system singular up to rounding
        assert_parity(ipx.approx_df(x, f, method, 0), O.approx_df(x, f, method, 0).astype(dtype), method)
    z = np.zeros((12, 2), dtype=dtype)
    xx = np.arange(12).astype(dtype)
    for method in O.CUBIC_METHODS:
        assert_parity(ipx.approx_df(xx, z, method, 0), O.approx_df(xx, z, method, 0).astype(dtype), method)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", ["cubic", "cardinal", "catmull-rom", "monotonic", "monotonic-0"])
def test_fused_slopes_pack_is_bitwise_the_table_chain(method, dtype):
    """ipx_slopes_pack_* (one pass f -> records) against the per-table chain
    fx, fy, fz <- f; fxy <- fx; fxz, fyz; fxyz <- fxy followed by ipx_pack_*: same bits, also at
    the one-sided ends, with duplicate knots, flat stretches and a trailing batch."""
    rng = np.random.default_rng(41)
    kw = {"c": 0.3} if method == "cardinal" else {}
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    for shape, batch in [((3, 3), ()), ((17, 9), ()), ((6, 5), (3,)), ((3, 3, 3), ()), ((9, 4, 13), ()),
                         ((5, 6, 7), (2,)), ((33, 18, 21), ())]:
        ndim = len(shape)
        kn = [knots_nonuniform(rng, n, 0, 2, dtype) for n in shape]
        if shape[0] > 8:
            kn[0][4] = kn[0][5]  # a zero-width interval
        f = rng.normal(size=shape + batch).astype(dtype)
        if shape[-1] > 8:
            f[..., 2:6] = f[..., 2:3]  # flat stretch along the last axis
        kd = [torch.as_tensor(k, device="cuda") for k in kn]
        fd = torch.as_tensor(f, device="cuda")
        nb = int(np.prod(batch, dtype=np.int64))
        got = api._slopes_pack_dev(ndim, kd, fd, nb, method, kw)
        assert got is not None and got.dtype == tdt
        tabs = {"f": fd}
        for name, src, ax in api.CHAIN[ndim]:
            tabs[name] = api._slopes_dev(kd[ax], tabs[src], method, ax, kw)
        want = api._pack_dev(ndim, [tabs[n] for n in api.TABLES[ndim]])
        g, w = got.cpu().numpy(), want.cpu().numpy()
        assert np.array_equal(g.view(np.uint8), w.view(np.uint8)), (method, shape, batch)
        # the class builds from the fused records and splits `.derivs` off them on demand
        cls = ipx.Interpolator2D if ndim == 2 else ipx.Interpolator3D
        obj = cls(*kd, fd, method=method, **kw)
        obj._build_records()  # (scalar cubic / cardinal tables evaluate from f and build these lazily)
        assert isinstance(obj._tabs, api._RecordTables)
        for name in api.TABLES[ndim][1:]:
            assert torch.equal(obj.derivs[name], tabs[name]), (method, shape, name)
        ref = (O.Interpolator2D if ndim == 2 else O.Interpolator3D)(*kn, f, method=method, **kw)
        q = [rng.uniform(-0.1, 2.1, 500).astype(dtype) for _ in range(ndim)]
        truth = None
        if dtype == np.float32:
            truth = (O.Interpolator2D if ndim == 2 else O.Interpolator3D)(
                *[k.astype(np.float64) for k in kn], f.astype(np.float64), method=method, **kw)(
                *[v.astype(np.float64) for v in q])
        assert_parity(obj(*[torch.as_tensor(v, device="cuda") for v in q]).cpu().numpy(), ref(*q),
                      f"{method} {shape}", truth)


def test_fused_slopes_pack_unsupported_cases_use_the_table_chain():
    rng = np.random.default_rng(42)
    x, y = np.linspace(0, 1, 2), np.linspace(0, 1, 7)
    f = rng.normal(size=(2, 7))
    obj = ipx.Interpolator2D(x, y, f, method="cubic")          # an axis with two knots
    obj._build_records()
    assert not isinstance(obj._tabs, api._RecordTables)
    obj = ipx.Interpolator2D(y, y, rng.normal(size=(7, 7)), method="akima")
    assert not isinstance(obj._tabs, api._RecordTables)
    lib = L.load()
    kd = [torch.as_tensor(y, device="cuda")] * 2
    narr = (C.c_int32 * 2)(7, 7)
    fd = torch.zeros(49, dtype=torch.float64, device="cuda")
    out = torch.zeros(49 * 4, dtype=torch.float64, device="cuda")
    call = lambda nd, m: lib.ipx_slopes_pack_f64(nd, api._ptr_array(kd), narr, api._ptr(fd), 1, m, None,
                                                 api._ptr(out), None)
    assert call(2, L.IPX_SLOPE_CUBIC) == L.IPX_OK
    assert call(2, L.IPX_SLOPE_AKIMA) == L.IPX_EUNSUPPORTED
    assert call(1, L.IPX_SLOPE_CUBIC) == L.IPX_EUNSUPPORTED
    assert call(2, 99) == L.IPX_EINVAL
    assert call(4, L.IPX_SLOPE_CUBIC) == L.IPX_EINVAL
    torch.cuda.synchronize()


@pytest.mark.parametrize("n", [2, 3, 4, 25])
def test_cubic2_boundary_conditions(n):
    rng = np.random.default_rng(5)
    x = knots_nonuniform(rng, n, 0, 3, np.float64)
    f = rng.normal(size=(n, 2, 3))
    v1 = rng.normal(size=(2, 3))
    v2 = rng.normal(size=(2, 3))
    kinds = ["not-a-knot", "natural", "clamped", (1, v1), (2, v2)]
    for bs, be in itertools.product(kinds, kinds):
        want = O.approx_df(x, f, "cubic2", 0, bc_type=(bs, be))
        got = ipx.approx_df(x, f, "cubic2", 0, bc_type=(bs, be))
        assert_parity(got, want, f"n={n} {bs if isinstance(bs, str) else bs[0]} {be if isinstance(be, str) else be[0]}")


def test_cubic2_boundary_values_complex_data():
    """Derivative-value boundary conditions on complex data: the values have the complex shape
    (not the (re, im)-pair shape the kernels see) and keep their imaginary part."""
    rng = np.random.default_rng(55)
    n = 9
    x = knots_nonuniform(rng, n, 0, 3, np.float64)
    f = rng.normal(size=(n, 4)) + 1j * rng.normal(size=(n, 4))
    v1 = rng.normal(size=4) + 1j * rng.normal(size=4)
    v2 = rng.normal(size=4) - 2j * rng.normal(size=4)
    for bc in (((1, v1), (2, v2)), ((2, v2), "natural"), ("clamped", (1, v1))):
        want = O.approx_df(x, f, "cubic2", 0, bc_type=bc)
        got = ipx.approx_df(x, f, "cubic2", 0, bc_type=bc)
        assert_parity(got, want, str(bc))
    with pytest.raises(ValueError):
        ipx.approx_df(x, f, "cubic2", 0, bc_type=((1, np.zeros((4, 2))), "natural"))


def test_approx_df_promotes_with_the_knots():
    """float32 data on float64 knots: the reference's dxi * df promotes to float64
    (_fd_derivs.py:84-90) and the knot spacing is never rounded to float32 first."""
    rng = np.random.default_rng(56)
    x = 1000.0 + np.cumsum(rng.uniform(0.5e-4, 1.5e-4, 50))  # spacing far below float32 resolution at 1e3
    f = rng.normal(size=50).astype(np.float32)
    for method in ("cubic", "cardinal", "monotonic", "akima", "cubic2"):
        want = O.approx_df(x, f.astype(np.float64), method, 0)
        got = ipx.approx_df(x, f, method, 0)
        assert got.dtype == np.float64
        # the reference differences the float32 data in float32 before it multiplies by the
        # float64 reciprocal spacing; the kernels widen the data first: equal to float32 rounding
        # of the differences (and nowhere near the error of differencing float32 KNOTS, ~1e-1 here)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max(), err_msg=method)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_akima_nan_and_inf_reach_the_global_threshold(dtype):
    """jnp.max(f12) propagates NaN (and inf), which sends EVERY node to the fall-back slope
    0.5 * (m[3:] + m[:-3]) (_fd_derivs.py:419-423): one bad value changes the slopes at all the
    finite nodes too, and the kernels follow."""
    rng = np.random.default_rng(57)
    x = knots_nonuniform(rng, 40, 0, 2, dtype)
    for bad in (np.nan, np.inf):
        f = rng.normal(size=(40, 3)).astype(dtype)
        f[17, 1] = bad
        with np.errstate(invalid="ignore"):
            want = O.approx_df(x, f, "akima", 0)
        got = ipx.approx_df(x, f, "akima", 0)
        fin = np.isfinite(want)
        assert fin.sum() > 60 and np.array_equal(np.isfinite(got), fin)
        rtol = RTOL[np.dtype(dtype)]
        np.testing.assert_allclose(got[fin], want[fin], rtol=rtol, atol=rtol * np.abs(want[fin]).max())


# ------------------------------------------------------------------ periodic knot preparation
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_periodic_knots_vs_oracle(dtype):
    rng = np.random.default_rng(6)
    P = 2 * np.pi
    cases = [
        np.linspace(0, P, 64, endpoint=False),          # sorted inside the period
        np.linspace(0, P, 200),                         # inclusive end: 2pi % 2pi == 0 -> tie across the wrap
        np.linspace(1.0, 1.0 + P, 50, endpoint=False),  # rotation
        np.linspace(-3.0, -3.0 + P, 81),                # rotation with a tie
        np.sort(rng.uniform(-20, 20, 97)),              # several periods: generic stable sort
        np.array([0.5, 0.5, 0.5, 7.0]),
    ]
    for x in cases:
        x = x.astype(dtype)
        xt = torch.from_numpy(x).cuda()
        xpad, perm = api._periodic_knots_dev(xt, P)
        _, want_x, want_p = O._make_periodic(np.zeros(1, dtype), x, P, 0, np.arange(len(x)))
        assert np.array_equal(xpad.cpu().numpy(), want_x)
        # the oracle's padded index vector is [perm[-1], perm..., perm[0]]
        assert np.array_equal(perm.cpu().numpy(), want_p[1:-1])


# ------------------------------------------------------------------ evaluation: randomised sweep
def make_problem(rng, ndim, dtype, batch_shape, uniform):
    ns = [int(rng.integers(5, 14)) for _ in range(ndim)]
    los = [float(rng.uniform(-2, 0)) for _ in range(ndim)]
    his = [lo + float(rng.uniform(1, 4)) for lo in los]
    if uniform:
        knots = [np.linspace(lo, hi, n).astype(dtype) for lo, hi, n in zip(los, his, ns)]
    else:
        knots = [knots_nonuniform(rng, n, lo, hi, dtype) for lo, hi, n in zip(los, his, ns)]
    f = rng.normal(size=(*ns, *batch_shape)).astype(dtype)
    nq = 513
    qs = []
    for lo, hi in zip(los, his):
        span = hi - lo
        q = rng.uniform(lo - 0.25 * span, hi + 0.25 * span, nq)
        q[:4] = [lo, hi, lo - 0.1, hi + 0.1]
        qs.append(q.astype(dtype))
    return knots, f, qs, los, his


EXTRAPS = {
    1: [False, True, 0.5, (True, False), (-1.0, True)],
    2: [False, True, 2.5, (False, True), ((True, 1.5), (False, True))],
    3: [False, True, -0.25, (3.0, False), ((True, True), (False, 2.0), (7.0, True))],
}


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", ALL_METHODS)
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_eval_sweep_vs_oracle(ndim, method, dtype):
    rng = np.random.default_rng(100 * ndim + len(method))
    ofun = (O.interp1d, O.interp2d, O.interp3d)[ndim - 1]
    gfun = (ipx.interp1d, ipx.interp2d, ipx.interp3d)[ndim - 1]
    ocls = (O.Interpolator1D, O.Interpolator2D, O.Interpolator3D)[ndim - 1]
    gcls = (ipx.Interpolator1D, ipx.Interpolator2D, ipx.Interpolator3D)[ndim - 1]
    trial = 0
    for batch_shape in [(), (3,), (2, 2)]:
        for uniform in (True, False):
            knots, f, qs, los, his = make_problem(rng, ndim, dtype, batch_shape, uniform)
            for extrap in EXTRAPS[ndim]:
                trial += 1
                if ndim == 1:
                    deriv = int(rng.integers(0, 5))
                    dcall = (deriv,)
                else:
                    deriv = tuple(int(v) for v in rng.integers(0, 4, ndim))
                    if trial % 3 == 0:
                        deriv = 0
                    dcall = (deriv,) * ndim if np.isscalar(deriv) else deriv
                if trial % 4 == 1:
                    per_axes = [bool(rng.integers(0, 2)) for _ in range(ndim)]
                    if not any(per_axes):
                        per_axes[0] = True
                    period = tuple((hi - lo) * 1.07 if p else None for lo, hi, p in zip(los, his, per_axes))
                    if ndim == 1:
                        period = period[0]
                else:
                    period = None
                kw = {"c": 0.25} if method == "cardinal" else {}
                what = f"nd={ndim} {method} B={batch_shape} uni={uniform} ex={extrap} d={deriv} per={period}"
                want = ofun(*qs, *knots, f, method, deriv, extrap, period, **kw)
                got = gfun(*qs, *knots, f, method, deriv, extrap, period, **kw)
                want_c = ocls(*knots, f, method, extrap, period, **kw)(*qs, *dcall)
                got_c = gcls(*knots, f, method, extrap, period, **kw)(*qs, *dcall)
                truth = truth_c = None
                if dtype == np.float32 and method != "nearest":
                    q64, k64 = [upcast(q) for q in qs], [upcast(k) for k in knots]
                    truth = ofun(*q64, *k64, upcast(f), method, deriv, extrap, period, **kw)
                    truth_c = ocls(*k64, upcast(f), method, extrap, period, **kw)(*q64, *dcall)
                assert_parity(got, want, "functional " + what, truth)
                assert_parity(got_c, want_c, "class " + what, truth_c)


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_eval_indices_and_layouts(ndim):
    """Indices from the fused kernel are bit-exact, and the AoS (packed) and SoA table
    layouts agree to the last few bits (same arithmetic, different loads)."""
    rng = np.random.default_rng(40 + ndim)
    knots, f, qs, los, his = make_problem(rng, ndim, np.float64, (2,), False)
    for q, k in zip(qs, knots):
        q[10] = np.nan
        q[11:11 + len(k)] = k[: len(q) - 11][: len(k)]
    res, idx = api._interp(ndim, qs, knots, f, "cubic", 0, True, None, {}, return_index=True)
    for ax in range(ndim):
        assert np.array_equal(idx[ax], O.bin_index(knots[ax], qs[ax]))
    cls = (ipx.Interpolator1D, ipx.Interpolator2D, ipx.Interpolator3D)[ndim - 1]
    got = cls(*knots, f, "cubic", True)(*qs)
    assert np.array_equal(np.isnan(got), np.isnan(res))
    m = ~np.isnan(res)
    # same arithmetic; only FMA contraction may differ between the two instantiations
    np.testing.assert_allclose(got[m], res[m], rtol=1e-13, atol=1e-13 * np.abs(res[m]).max())


def test_nearest_1d_ties_and_far_queries():
    x = np.array([0.0, 1.0, 1.0, 2.0, 4.0])
    f = np.array([10.0, 11.0, 12.0, 13.0, 14.0])
    q = np.array([0.5, 1.0, 1.5, 3.0, -5.0, 9.0, 1e300, -1e300, np.nan, 2.0, 0.0])
    want = O.interp1d(q, x, f, "nearest", 0, True)
    got = ipx.interp1d(q, x, f, "nearest", 0, True)
    assert np.array_equal(got, want)
    xf = x.astype(np.float32)
    qf = np.array([0.5, 1e30, -1e30, 3.0000001, 1.0], dtype=np.float32)
    assert np.array_equal(ipx.interp1d(qf, xf, f.astype(np.float32), "nearest", 0, True),
                          O.interp1d(qf, xf, f.astype(np.float32), "nearest", 0, True))


# ------------------------------------------------------------------ the reference's own tests, restated
TOL_1D = {
    "nearest": (1e-2, 1e-1), "linear": (1e-4, 1e-3), "cubic": (1e-6, 1e-5), "cubic2": (1e-6, 1e-5),
    "cardinal": (1e-6, 1e-5), "catmull-rom": (1e-6, 1e-5), "monotonic": (1e-4, 1e-3),
    "monotonic-0": (1e-4, 1.5e-2), "akima": (1e-6, 2e-5),
}


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("scalar", [False, True])
def test_reference_test_interp1d(dtype, scalar):
    # tests/test_interpolate.py:26-82
    xp = np.real(np.linspace(0, 2 * np.pi, 100).astype(dtype))
    cplx = np.iscomplexobj(np.zeros(1, dtype))
    f = (lambda x: np.sin(x) + 1j * np.sin(x)) if cplx else (lambda x: np.sin(x))
    fp = f(xp).astype(dtype)
    x = 0.0 if scalar else np.real(np.linspace(0, 2 * np.pi, 10000).astype(dtype))
    for interp in (ipx.interp1d, lambda xq, *a, **k: ipx.Interpolator1D(*a, **k)(xq)):
        for m, (rtol, atol) in TOL_1D.items():
            fq = interp(x, xp, fp, method=m)
            assert fq.dtype == np.result_type(x, xp, fp), m
            np.testing.assert_allclose(fq, f(x), rtol=rtol, atol=atol, err_msg=m)
            assert_parity(fq, O.interp1d(x, xp, fp, method=m), m)


def test_reference_test_interp1d_extrap_periodic_monotonic():
    # tests/test_interpolate.py:117-149
    xp = np.linspace(0, 2 * np.pi, 200)
    x = np.linspace(-1, 2 * np.pi + 1, 10000)
    fp = np.sin(xp)
    fq = ipx.interp1d(x, xp, fp, method="cubic", extrap=False)
    assert np.isnan(fq[0]) and np.isnan(fq[-1])
    fq = ipx.interp1d(x, xp, fp, method="cubic", extrap=True)
    assert not np.isnan(fq[0]) and not np.isnan(fq[-1])
    fq = ipx.interp1d(x, xp, fp, method="cubic", period=2 * np.pi)
    np.testing.assert_allclose(fq, np.sin(x), rtol=1e-6, atol=1e-2)
    assert_parity(fq, O.interp1d(x, xp, fp, method="cubic", period=2 * np.pi))
    xk = np.linspace(-4, 5, 10)
    f = np.heaviside(xk - 1.5, 0) + 0.1 * xk
    xq = np.linspace(-4, 5, 1000)
    dfc = ipx.interp1d(xq, xk, f, derivative=1, method="cubic")
    dfm = ipx.interp1d(xq, xk, f, derivative=1, method="monotonic")
    dfm0 = ipx.interp1d(xq, xk, f, derivative=1, method="monotonic-0")
    assert dfc.min() < 0 and dfm.min() > 0 and dfm0.min() >= 0
    np.testing.assert_allclose(dfm0[np.array([0, -1])], 0, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("scalar", [False, True])
def test_reference_test_interp2d(dtype, scalar):
    # tests/test_interpolate.py:155-235
    cplx = np.iscomplexobj(np.zeros(1, dtype))
    xp = np.real(np.linspace(0, 3 * np.pi, 99).astype(dtype))
    yp = np.real(np.linspace(0, 2 * np.pi, 40).astype(dtype))
    xxp, yyp = np.meshgrid(xp, yp, indexing="ij")
    f = (lambda x, y: np.sin(x) * np.cos(y) * (1 + 1j)) if cplx else (lambda x, y: np.sin(x) * np.cos(y))
    fp = f(xxp, yyp).astype(dtype)
    if scalar:
        x, y = 0.0, 0.0
    else:
        x = np.real(np.linspace(0, 3 * np.pi, 1000).astype(dtype))
        y = np.real(np.linspace(0, 2 * np.pi, 1000).astype(dtype))
    per = (2 * np.pi, 2 * np.pi)
    tol = {"nearest": (1e-2, 1), "linear": (1e-4, 1e-2), "cubic": (1e-5, 2e-3), "cubic2": (1e-5, 2e-3),
           "catmull-rom": (1e-5, 2e-3), "cardinal": (1e-5, 2e-3), "akima": (1e-5, 2e-3),
           "monotonic": (1e-2, 1e-3), "monotonic-0": (1e-2, 2e-2)}
    for form, interp in (("fn", ipx.interp2d), ("cls", lambda xq, yq, *a, **k: ipx.Interpolator2D(*a, **k)(xq, yq))):
        ointerp = O.interp2d if form == "fn" else (lambda xq, yq, *a, **k: O.Interpolator2D(*a, **k)(xq, yq))
        for m, (rtol, atol) in tol.items():
            fq = interp(x, y, xp, yp, fp, method=m, period=per)
            assert fq.dtype == np.result_type(x, y, xp, yp, fp)
            np.testing.assert_allclose(fq, f(x, y), rtol=rtol, atol=atol, err_msg=m)
            assert_parity(fq, ointerp(x, y, xp, yp, fp, method=m, period=per), f"{form} {m}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("scalar", [False, True])
def test_reference_test_interp3d(dtype, scalar):
    # tests/test_interpolate.py:262-337
    cplx = np.iscomplexobj(np.zeros(1, dtype))
    xp = np.real(np.linspace(0, np.pi, 20).astype(dtype))
    yp = np.real(np.linspace(0, 2 * np.pi, 30).astype(dtype))
    zp = np.real(np.linspace(0, 3, 25).astype(dtype))
    xxp, yyp, zzp = np.meshgrid(xp, yp, zp, indexing="ij")
    g = lambda x, y, z: np.sin(x) * np.cos(y) * z**2
    f = (lambda x, y, z: g(x, y, z) * (1 + 1j)) if cplx else g
    fp = f(xxp, yyp, zzp).astype(dtype)
    if scalar:
        x, y, z = 0.0, 0.0, 0.0
    else:
        x = np.real(np.linspace(0, np.pi, 1000).astype(dtype))
        y = np.real(np.linspace(0, 2 * np.pi, 1000).astype(dtype))
        z = np.real(np.linspace(0, 3, 1000).astype(dtype))
    tol = {"nearest": (1e-2, 1), "linear": (1e-3, 1e-1), "cubic": (1e-5, 5.5e-3), "cubic2": (1e-5, 5.5e-3),
           "catmull-rom": (1e-5, 5.5e-3), "cardinal": (1e-5, 5.5e-3), "monotonic": (1e-2, 1e-2),
           "monotonic-0": (1e-2, 0.3)}
    for form, interp in (("fn", ipx.interp3d),
                         ("cls", lambda a, b, c, *r, **k: ipx.Interpolator3D(*r, **k)(a, b, c))):
        ointerp = O.interp3d if form == "fn" else (lambda a, b, c, *r, **k: O.Interpolator3D(*r, **k)(a, b, c))
        for m, (rtol, atol) in tol.items():
            fq = interp(x, y, z, xp, yp, zp, fp, method=m)
            assert fq.dtype == np.result_type(x, y, z, xp, yp, zp, fp)
            np.testing.assert_allclose(fq, f(x, y, z), rtol=rtol, atol=atol, err_msg=m)
            assert_parity(fq, ointerp(x, y, z, xp, yp, zp, fp, method=m), f"{form} {m}")


def test_reference_test_vector_valued_and_extrap_float():
    # tests/test_interpolate.py:85-114, 238-256, 340-369, 640-649
    xp = np.linspace(0, 2 * np.pi, 100)
    x = np.linspace(0, 2 * np.pi, 300)[10:-10]
    f = lambda x: np.array([np.sin(x), np.cos(x)])
    fp = f(xp).T
    for m in ("nearest", "linear", "cubic", "cubic2", "cardinal", "catmull-rom", "monotonic", "monotonic-0"):
        rtol, atol = TOL_1D[m]
        fq = ipx.interp1d(x, xp, fp, method=m)
        np.testing.assert_allclose(fq, f(x).T, rtol=rtol, atol=max(atol, 1e-2 if m == "monotonic-0" else 0))
        assert_parity(fq, O.interp1d(x, xp, fp, method=m), m)
    xk = np.linspace(0, 10, 10)
    yk = np.linspace(0, 8, 8)
    z = np.zeros((10, 8)) + 1.0
    ip = ipx.Interpolator2D(xk, yk, z, extrap=0.0)
    np.testing.assert_allclose(ip(4.5, 5.3), 1.0)
    np.testing.assert_allclose(ip(-4.5, 5.3), 0.0)
    np.testing.assert_allclose(ip(4.5, -5.3), 0.0)
    # 3-D vector valued
    xs = np.linspace(0, np.pi, 1000); ys = np.linspace(0, 2 * np.pi, 1000); zs = np.linspace(0, 3, 1000)
    xp = np.linspace(0, np.pi, 20); yp = np.linspace(0, 2 * np.pi, 30); zp = np.linspace(0, 3, 25)
    xxp, yyp, zzp = np.meshgrid(xp, yp, zp, indexing="ij")
    g = lambda x, y, z: np.array([np.sin(x) * np.cos(y) * z**2, 0.1 * (x + y - z)])
    fp = g(xxp.T, yyp.T, zzp.T).T
    for m, (rtol, atol) in {"nearest": (1e-2, 1), "linear": (1e-3, 1e-1), "cubic": (1e-5, 5e-3),
                            "monotonic": (1e-2, 1e-2)}.items():
        fq = ipx.interp3d(xs, ys, zs, xp, yp, zp, fp, method=m)
        np.testing.assert_allclose(fq, g(xs, ys, zs).T, rtol=rtol, atol=atol)
        assert_parity(fq, O.interp3d(xs, ys, zs, xp, yp, zp, fp, method=m), m)


# ------------------------------------------------------------------ "grad wrt xq" = derivative+1
def test_derivative_kernel_matches_finite_differences():
    """The jax.ffi binding's JVP for xq re-enters the kernel with derivative+1
    (INTEGRATION.md); check that this is the derivative of the derivative-0 kernel."""
    rng = np.random.default_rng(9)
    xp = np.linspace(0, 1, 12); yp = np.linspace(0, 2, 9); zp = np.linspace(-1, 1, 10)
    f = rng.normal(size=(12, 9, 10))
    ip = ipx.Interpolator3D(xp, yp, zp, f, "cubic")
    q = [rng.uniform(a[1], a[-2], 200) for a in (xp, yp, zp)]
    h = 1e-6
    for ax in range(3):
        qp = [v.copy() for v in q]; qm = [v.copy() for v in q]
        qp[ax] += h; qm[ax] -= h
        fd = (ip(*qp) - ip(*qm)) / (2 * h)
        d = [0, 0, 0]; d[ax] = 1
        np.testing.assert_allclose(ip(*q, *d), fd, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ C ABI specifics
def test_abi_device_params_graph_capture_and_errors():
    lib = L.load()
    dev = torch.device("cuda")
    rng = np.random.default_rng(10)
    n = 64
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    f = torch.from_numpy(rng.normal(size=n)).to(dev)
    fx = api._slopes_dev(x, f, "cubic", 0, {})
    q = torch.from_numpy(rng.uniform(-0.1, 1.1, 4096)).to(dev)
    out_h = torch.empty(4096, dtype=torch.float64, device=dev)
    out_d = torch.empty_like(out_h)
    out_g = torch.empty_like(out_h)
    p = api._make_params(1, (1,), (False, 3.0), (None,))
    tabs = api._ptr_array([f, fx])
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.ipx_eval1d_f64(q.data_ptr(), 4096, x.data_ptr(), None, n, tabs, L.IPX_SOA, 1, L.IPX_CUBIC, 0,
                            C.byref(p), 0, out_h.data_ptr(), None, stream)
    assert rc == L.IPX_OK
    # traced operands from DEVICE memory (what an XLA custom call would hand over)
    pbytes = torch.frombuffer(bytearray(bytes(p)), dtype=torch.uint8).to(dev)
    rc = lib.ipx_eval1d_f64(q.data_ptr(), 4096, x.data_ptr(), None, n, tabs, L.IPX_SOA, 1, L.IPX_CUBIC, 0,
                            C.cast(C.c_void_p(pbytes.data_ptr()), C.POINTER(L.Params)), 1,
                            out_d.data_ptr(), None, stream)
    assert rc == L.IPX_OK
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(out_h, nan=-7.0), torch.nan_to_num(out_d, nan=-7.0))
    want = O.interp1d(q.cpu().numpy(), x.cpu().numpy(), f.cpu().numpy(), "cubic", 1, (False, 3.0))
    assert_parity(out_h.cpu().numpy(), want)
    # the call only enqueues on the given stream: it can be captured in a CUDA graph ("jit")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            rc = lib.ipx_eval1d_f64(q.data_ptr(), 4096, x.data_ptr(), None, n, tabs, L.IPX_SOA, 1,
                                    L.IPX_CUBIC, 0, C.byref(p), 0, out_g.data_ptr(), None,
                                    torch.cuda.current_stream().cuda_stream)
            assert rc == L.IPX_OK
    out_g.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(out_h, nan=-7.0), torch.nan_to_num(out_g, nan=-7.0))
    # error codes, not aborts
    assert lib.ipx_eval1d_f64(q.data_ptr(), 4096, x.data_ptr(), None, 1, tabs, L.IPX_SOA, 1, L.IPX_CUBIC, 0,
                              C.byref(p), 0, out_h.data_ptr(), None, stream) == L.IPX_EINVAL
    assert lib.ipx_eval1d_f64(q.data_ptr(), 4096, x.data_ptr(), None, n, tabs, 7, 1, L.IPX_CUBIC, 0,
                              C.byref(p), 0, out_h.data_ptr(), None, stream) == L.IPX_EINVAL
    assert lib.ipx_eval1d_f64(q.data_ptr(), 0, x.data_ptr(), None, n, tabs, L.IPX_SOA, 1, L.IPX_CUBIC, 0,
                              C.byref(p), 0, None, None, stream) == L.IPX_OK  # empty input


def test_torch_inputs_stay_on_device_and_empty_queries():
    dev = torch.device("cuda")
    xp = torch.linspace(0, 1, 20, dtype=torch.float64, device=dev)
    fp = torch.sin(xp)
    q = torch.rand(1000, dtype=torch.float64, device=dev)
    out = ipx.interp1d(q, xp, fp)
    assert isinstance(out, torch.Tensor) and out.is_cuda and out.shape == (1000,)
    assert_parity(out.cpu().numpy(), O.interp1d(q.cpu().numpy(), xp.cpu().numpy(), fp.cpu().numpy()))
    empty = ipx.interp1d(np.zeros(0), xp.cpu().numpy(), fp.cpu().numpy())
    assert empty.shape == (0,)
    e3 = ipx.interp3d(np.zeros(0), np.zeros(0), np.zeros(0), np.arange(3.), np.arange(3.), np.arange(3.),
                      np.zeros((3, 3, 3, 2)))
    assert e3.shape == (0, 2)


# ------------------------------------------------------------------ full-size properties (C3 / C4 shapes)
def test_full_size_properties_c4():
    """256^3 tricubic, period on x and y, extrap on z, 2^24 queries: size-independent
    properties that need no oracle -- exactness at the knots, constants reproduced,
    linearity in the table, invariance under query permutation, periodic images agree,
    plus oracle parity on a 2^14-query subsample."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1234 + 4)
    n = 256
    P = 2 * np.pi
    x = torch.arange(n, dtype=torch.float64, device=dev) * (P / n)
    z = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    f1 = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
    f2 = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
    period = (P, P, None)
    extrap = ((True, True), (True, True), (True, 0.0))
    ip1 = ipx.Interpolator3D(x, x, z, f1, "cubic", extrap, period)
    nq = 1 << 24
    xq = (torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P
    yq = (torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P
    zq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.1 - 0.05
    r1 = ip1(xq, yq, zq)
    # fill value above the open axis, polynomial continuation below it
    assert torch.all(r1[zq > 1.0] == 0.0) and torch.isfinite(r1).all()
    # permutation invariance (bitwise): each query is independent of its neighbours
    perm = torch.randperm(nq, device=dev, generator=g)
    assert torch.equal(ip1(xq[perm], yq[perm], zq[perm]), r1[perm])
    # periodic images: q and q + P land in the same cell up to the rounding of the wrap
    rimg = ip1(xq + P, yq - P, zq)
    assert torch.allclose(rimg, r1, rtol=0, atol=1e-9 * float(f1.abs().max()) * n)
    # linearity in the table (slopes are linear for method="cubic")
    ip2 = ipx.Interpolator3D(x, x, z, f2, "cubic", extrap, period)
    ip12 = ipx.Interpolator3D(x, x, z, 2.0 * f1 - 3.0 * f2, "cubic", extrap, period)
    lin = 2.0 * r1 - 3.0 * ip2(xq, yq, zq)
    assert torch.allclose(ip12(xq, yq, zq), lin, rtol=1e-11, atol=1e-11 * float(lin.abs().max()))
    # constants are reproduced exactly up to rounding; knots are reproduced
    ipc = ipx.Interpolator3D(x, x, z, torch.full_like(f1, 3.25), "cubic", True, period)
    inside = (zq >= 0) & (zq <= 1)  # (far outside, the weights cancel from ~|t|^3)
    rc = ipc(xq, yq, zq)
    assert torch.allclose(rc[inside], torch.full_like(rc[inside], 3.25), rtol=1e-13, atol=0)
    ii = torch.randint(0, n, (4096,), device=dev, generator=g)
    jj = torch.randint(0, n, (4096,), device=dev, generator=g)
    kk = torch.randint(0, n, (4096,), device=dev, generator=g)
    atk = ip1(x[ii], x[jj], z[kk])
    assert torch.allclose(atk, f1[ii, jj, kk], rtol=1e-13, atol=1e-13)
    # oracle parity on a subsample (class form: slopes on the un-padded table)
    m = 1 << 14
    want = O.Interpolator3D(x.cpu().numpy(), x.cpu().numpy(), z.cpu().numpy(), f1.cpu().numpy(), "cubic",
                            extrap, period)(xq[:m].cpu().numpy(), yq[:m].cpu().numpy(), zq[:m].cpu().numpy())
    assert_parity(r1[:m].cpu().numpy(), want)


def test_full_size_properties_c3():
    """2048^2 bicubic, 2^24 random queries, derivative (0,0) and (1,0), f64."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1234 + 3)
    n = 2048
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    f = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
    ip = ipx.Interpolator2D(x, x, f, "cubic", False)
    nq = 1 << 24
    xq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
    yq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
    r0 = ip(xq, yq)
    r1 = ip(xq, yq, 1, 0)
    outside = (xq < 0) | (xq > 1) | (yq < 0) | (yq > 1)
    assert torch.equal(torch.isnan(r0), outside) and torch.equal(torch.isnan(r1), outside)
    m = 1 << 14
    o = O.Interpolator2D(x.cpu().numpy(), x.cpu().numpy(), f.cpu().numpy(), "cubic", False)
    xs, ys = xq[:m].cpu().numpy(), yq[:m].cpu().numpy()
    assert_parity(r0[:m].cpu().numpy(), o(xs, ys))
    assert_parity(r1[:m].cpu().numpy(), o(xs, ys, 1, 0))


# ------------------------------------------------------------------ host-resident queries
def test_host_streaming_matches_device_path():
    """NumPy / CPU-tensor queries are streamed through the GPU in chunks over several CUDA
    streams; the result must be bit-identical to the all-on-device call, for pageable and
    pinned buffers, with and without a caller-provided `out`."""
    from interpax_b200 import api

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(77)
    n = 48
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    f = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
    ip = ipx.Interpolator3D(x, x, x, f, "cubic", True)
    nq = 2 * api._HOST_CHUNK + 12345  # three chunks, odd tail
    qd = [torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.1 - 0.05 for _ in range(3)]
    want = ip(*qd)
    assert want.is_cuda
    qh = [q.cpu() for q in qd]
    got = ip(*qh)                                   # pageable CPU tensors -> CPU tensor
    assert isinstance(got, torch.Tensor) and not got.is_cuda
    assert torch.equal(got, want.cpu())
    qp = [q.pin_memory() for q in qh]
    out = torch.empty(nq, dtype=torch.float64, pin_memory=True)
    ret = ip(*qp, out=out)                          # pinned in, pinned out
    assert ret.data_ptr() == out.data_ptr() and torch.equal(out, want.cpu())
    xn = x.cpu().numpy()                            # NumPy everywhere -> NumPy out
    ipn = ipx.Interpolator3D(xn, xn, xn, f.cpu().numpy(), "cubic", True)
    got_np = ipn(*[q.numpy() for q in qh])
    assert isinstance(got_np, np.ndarray) and np.array_equal(got_np, want.cpu().numpy())
    assert api.launches_per_host_call(nq) == 3
    # batch > 1 through the same path
    fb = torch.randn((n, n, 3), dtype=torch.float64, device=dev, generator=g)
    ip2 = ipx.Interpolator2D(x, x, fb, "cubic", True)
    m = api._HOST_CHUNK + 7
    assert torch.equal(ip2(qh[0][:m], qh[1][:m]), ip2(qd[0][:m], qd[1][:m]).cpu())
    with pytest.raises(ValueError):
        ip(*qd, out=out)                            # device queries need a device `out`


# ------------------------------------------------------------------ wide vector-valued 1-D tables (C2 shape)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("B", [32, 64, 70, 130])
def test_wide_batch_1d_rows_kernel(B, dtype):
    """f of shape (Nx, B) with B >= 32 takes the row-streaming kernel (lanes = columns, rows
    reused while the interval does not change): functional (SoA) and class (AoS) forms, linear
    and cubic, sorted / random / clustered queries, extrapolation fills, periodic wrap,
    derivatives, a query count that is not a multiple of the warp size."""
    rng = np.random.default_rng(B)
    n = 57
    x = knots_nonuniform(rng, n, -1.0, 2.0, dtype)
    f = rng.normal(size=(n, B)).astype(dtype)
    nq = 1000 + 13
    q_rand = rng.uniform(-1.3, 2.3, nq).astype(dtype)
    q_sort = np.sort(q_rand)
    q_clus = np.repeat(rng.uniform(-1, 2, 40), 26)[:nq].astype(dtype)
    q_clus[5] = np.nan
    up = lambda v: v.astype(np.float64) if dtype == np.float32 else None
    for q in (q_rand, q_sort, q_clus):
        for method, deriv, extrap, period in [
            ("cubic", 0, False, None), ("cubic", 1, (True, 0.5), None), ("linear", 0, (2.0, False), None),
            ("linear", 1, True, None), ("monotonic", 2, True, None), ("cubic", 0, False, 3.3),
            ("akima", 0, True, None), ("cubic2", 3, False, None),
        ]:
            what = f"B={B} {method} d={deriv} ex={extrap} per={period}"
            want = O.interp1d(q, x, f, method, deriv, extrap, period)
            got = ipx.interp1d(q, x, f, method, deriv, extrap, period)
            truth = None
            if dtype == np.float32:
                truth = O.interp1d(up(q), up(x), up(f), method, deriv, extrap, period)
            assert_parity(got, want, "fn " + what, truth)
            want_c = O.Interpolator1D(x, f, method, extrap, period)(q, deriv)
            got_c = ipx.Interpolator1D(x, f, method, extrap, period)(q, deriv)
            truth_c = None
            if dtype == np.float32:
                truth_c = O.Interpolator1D(up(x), up(f), method, extrap, period)(up(q), deriv)
            assert_parity(got_c, want_c, "cls " + what, truth_c)
    res, idx = api._interp(1, (q_rand,), (x,), f, "cubic", 0, True, None, {}, return_index=True)
    assert np.array_equal(idx[0], O.bin_index(x, q_rand))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cubic2_long_axis_windowed_solve(dtype):
    """Axes longer than 1024 knots use the windowed Thomas solve (chunks of 256 rows entered
    through 64-row halos); it must agree with the sequential sweep of the oracle to rounding,
    for every boundary-condition kind, uniform and non-uniform knots, any table layout."""
    rng = np.random.default_rng(11)
    for n, shape, axis in [(1500, (1500,), 0), (4099, (4099, 3), 0), (2048, (2, 2048), 1), (1025, (2, 1025, 3), 1)]:
        for uniform in (True, False):
            x = (np.linspace(0, 5, n) if uniform else knots_nonuniform(rng, n, 0, 5, np.float64)).astype(dtype)
            f = np.cumsum(rng.normal(size=shape), axis=axis).astype(dtype) * dtype(0.1)
            vshape = tuple(sh for k, sh in enumerate(shape) if k != axis)
            v1 = rng.normal(size=vshape).astype(dtype)
            for bc in ("not-a-knot", "natural", "clamped", ((1, v1), "not-a-knot"), ("natural", (2, v1))):
                want = O.approx_df(x, f, "cubic2", axis, bc_type=bc)
                got = ipx.approx_df(x, f, "cubic2", axis, bc_type=bc)
                truth = None
                if dtype == np.float32:
                    bc64 = bc if isinstance(bc, str) else tuple(b if isinstance(b, str) else (b[0], b[1].astype(np.float64)) for b in bc)
                    truth = O.approx_df(x.astype(np.float64), f.astype(np.float64), "cubic2", axis, bc_type=bc64)
                assert_parity(got, want.astype(dtype), f"n={n} {shape} uniform={uniform} bc={bc if isinstance(bc, str) else 'mixed'}", truth)


# ------------------------------------------------------------------ spatial pre-sort of the queries
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_presort_is_bitwise_identical(ndim):
    """presort=True (cell-order radix sort -> gather -> evaluate -> scatter) must give exactly
    the results of the plain call, in the caller's order, for every method class, with
    periodic axes, extrapolation fills, NaN queries, batches, and from host memory (scalar cubic
    tables: to rounding, see below)."""
    rng = np.random.default_rng(60 + ndim)
    dev = torch.device("cuda")
    ns = [37, 29, 23][:ndim]
    knots = [knots_nonuniform(rng, n, 0.0, 2.0, np.float64) for n in ns]
    nq = 200_003
    qs = [rng.uniform(-0.3, 2.3, nq) for _ in range(ndim)]
    qs[0][7] = np.nan
    cls = (ipx.Interpolator1D, ipx.Interpolator2D, ipx.Interpolator3D)[ndim - 1]
    fun = (ipx.interp1d, ipx.interp2d, ipx.interp3d)[ndim - 1]
    for batch_shape in ((), (3,)):
        f = rng.normal(size=(*ns, *batch_shape))
        for method, extrap, period in (("cubic", (True, 0.5), None), ("linear", False, None),
                                      ("nearest", True, None), ("monotonic", False, 2.2)):
            per = period if ndim == 1 or period is None else (period,) + (None,) * (ndim - 1)
            ip = cls(*knots, f, method, extrap, per)
            a = ip(*qs)
            b = ip(*qs, presort=True)
            assert np.array_equal(np.isnan(a), np.isnan(b))
            m = ~np.isnan(a)
            if method == "cubic" and batch_shape == ():
                # scalar cubic tables have two forms (evaluation from f, packed records) and the
                # device picks one per stream (random order vs cell-sorted): equal to rounding
                np.testing.assert_allclose(a[m], b[m], rtol=1e-13, atol=1e-13 * np.abs(a[m]).max())
            else:
                assert np.array_equal(a[m], b[m]), (method, batch_shape)
            c = fun(*qs, *knots, f, method, 0, extrap, per, presort=True)
            assert_parity(c, fun(*qs, *knots, f, method, 0, extrap, per), method)
    # the order really is the cell order
    from interpax_b200 import api
    qd = [torch.from_numpy(q).to(dev) for q in qs]
    kd = [torch.from_numpy(k).to(dev) for k in knots]
    params = api._make_params(ndim, (0,) * ndim, (True,) * (2 * ndim), (None,) * ndim)
    order = api._cell_order_dev(ndim, qd, kd, ns, 0, params).cpu().numpy().astype(np.int64)
    assert np.array_equal(np.sort(order), np.arange(nq))
    key = np.zeros(nq, dtype=np.int64)
    for ax in range(ndim):
        key = key * (ns[ax] - 1) + (O.bin_index(knots[ax], qs[ax]) - 1)
    assert np.array_equal(order, np.argsort(key, kind="stable"))


def test_cuda_graph_capture_of_the_class_call():
    """"jit" for the headline path: Interpolator3D.__call__ (tile kernel, periodic axes, packed
    tables) captured in a CUDA graph and replayed on new query values."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(5)
    n = 40
    P = 2 * np.pi
    x = torch.arange(n, device=dev, dtype=torch.float64) * (P / n)
    z = torch.linspace(0, 1, n, device=dev, dtype=torch.float64)
    f = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
    ip = ipx.Interpolator3D(x, x, z, f, "cubic", ((True, True), (True, True), (True, 0.0)), (P, P, None))
    nq = 50_000
    q = [(torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P for _ in range(2)]
    q.append(torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.1 - 0.05)
    out = torch.empty(nq, dtype=torch.float64, device=dev)
    ip(*q, out=out)  # warm-up: builds the periodic-knot cache (one host sync), loads the kernel
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            ip(*q, 1, 0, 0, out=out)
    for _ in range(2):
        for t in q[:2]:
            t.copy_((torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P)
        graph.replay()
        torch.cuda.synchronize()
        want = ip(*q, 1, 0, 0)
        assert torch.equal(out, want)


def test_more_than_2_to_31_queries():
    """64-bit indexing: 2^31 + 37 queries through the tile kernel (1-D f64 cubic, class form)
    and the generic kernel (1-D f32 linear); checked at the far end and at random positions
    against the same calls on small slices."""
    dev = torch.device("cuda")
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2**30:
        pytest.skip("needs ~45 GB of free device memory")
    nq = 2**31 + 37
    n = 1000
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    f = torch.sin(7 * x)
    ip = ipx.Interpolator1D(x, f, "cubic", True)
    q = torch.empty(nq, dtype=torch.float64, device=dev)
    q[:] = torch.linspace(-0.01, 1.01, 1 << 20, dtype=torch.float64, device=dev).repeat(nq // (1 << 20) + 1)[:nq]
    out = torch.empty(nq, dtype=torch.float64, device=dev)
    ip(q, out=out)
    for lo in (0, 2**31 - 1000, nq - 4096):
        assert torch.equal(out[lo:lo + 4096], ip(q[lo:lo + 4096].clone()))
    del out
    q32 = q.to(torch.float32)
    del q
    lin = ipx.interp1d(q32, x.to(torch.float32), f.to(torch.float32), "linear", 0, True)
    for lo in (0, 2**31 + 5, nq - 4096):
        want = ipx.interp1d(q32[lo:lo + 4096].clone(), x.to(torch.float32), f.to(torch.float32), "linear", 0, True)
        assert torch.equal(lin[lo:lo + 4096][:want.numel()], want)


# ------------------------------------------------------------------ reverse mode w.r.t. the tables
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_vjp_tables_is_the_transpose_of_the_evaluation(ndim):
    """<cot, eval(tables + dT) - eval(tables)> == sum_tables <grad, dT>: the evaluation is linear
    in its tables (away from extrapolation fills), so the identity is exact up to rounding.
    Checked for the cubic family (all 2^d tables) and linear, with periodic axes, fills,
    derivatives and a batch, and against the oracle evaluating the perturbed tables."""
    rng = np.random.default_rng(80 + ndim)
    ns = [11, 9, 8][:ndim]
    names = {1: ("f", "fx"), 2: ("f", "fx", "fy", "fxy"), 3: ("f", "fx", "fy", "fz", "fxy", "fxz", "fyz", "fxyz")}[ndim]
    knots = [knots_nonuniform(rng, n, 0.0, 2.0, np.float64) for n in ns]
    nq = 3000
    qs = [rng.uniform(-0.2, 2.2, nq) for _ in range(ndim)]
    ofun = (O.interp1d, O.interp2d, O.interp3d)[ndim - 1]
    for batch_shape in ((), (2,)):
        fshape = (*ns, *batch_shape)
        for method, deriv, extrap, period in (("cubic", 0, True, None), ("cubic", 1, (True, 0.5), None),
                                              ("cubic", 0, False, 2.3), ("linear", 0, False, None),
                                              ("linear", 1, True, None)):
            per = period if (ndim == 1 or period is None) else (period,) + (None,) * (ndim - 1)
            tabs = {nm: rng.normal(size=fshape) for nm in names}
            dtab = {nm: rng.normal(size=fshape) for nm in names}
            cot = rng.normal(size=(nq, *batch_shape))
            grads = ipx.vjp_tables(qs, knots, fshape, cot, method, deriv, extrap, per)
            if method == "linear":
                assert set(grads) == {"f"}
                base = ofun(*qs, *knots, tabs["f"], method, deriv, extrap, per)
                pert = ofun(*qs, *knots, tabs["f"] + dtab["f"], method, deriv, extrap, per)
                rhs = float(np.sum(grads["f"] * dtab["f"]))
            else:
                assert set(grads) == set(names)
                kw = {nm: tabs[nm] for nm in names[1:]}
                kw2 = {nm: tabs[nm] + dtab[nm] for nm in names[1:]}
                base = ofun(*qs, *knots, tabs["f"], method, deriv, extrap, per, **kw)
                pert = ofun(*qs, *knots, tabs["f"] + dtab["f"], method, deriv, extrap, per, **kw2)
                rhs = float(sum(np.sum(grads[nm] * dtab[nm]) for nm in names))
            d = pert - base
            d = np.where(np.isfinite(d), d, 0.0)  # NaN fills (extrap=False) carry no gradient
            lhs = float(np.sum(cot * d))
            scale = float(np.sum(np.abs(cot) * np.abs(d))) + 1e-300
            assert abs(lhs - rhs) <= 1e-11 * scale, (method, deriv, extrap, per, lhs, rhs)


def test_full_size_properties_f32_3d():
    """float32 tile kernel (three resident CTAs per SM, 32-byte records) on a 128^3 periodic
    table with 2^22 queries: permutation invariance (bitwise), cell-sorted == random order,
    presort == plain, oracle parity on a subsample."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(99)
    n = 128
    P = 2 * np.pi
    x = (torch.arange(n, dtype=torch.float64, device=dev) * (P / n)).to(torch.float32)
    z = torch.linspace(0, 1, n, dtype=torch.float32, device=dev)
    f = torch.randn((n, n, n), dtype=torch.float32, device=dev, generator=g)
    period = (P, P, None)
    extrap = ((True, True), (True, True), (True, 0.0))
    ip = ipx.Interpolator3D(x, x, z, f, "cubic", extrap, period)
    nq = 1 << 22
    q = [(torch.rand(nq, dtype=torch.float32, device=dev, generator=g) * 3 - 1) * P for _ in range(2)]
    q.append(torch.rand(nq, dtype=torch.float32, device=dev, generator=g) * 1.1 - 0.05)
    r = ip(*q)
    assert r.dtype == torch.float32 and torch.isfinite(r).all()
    perm = torch.randperm(nq, device=dev, generator=g)
    assert torch.equal(ip(*[a[perm] for a in q]), r[perm])
    h = P / n
    key = ((torch.remainder(q[0], P) / h).floor().clamp(0, n - 1).long() * n
           + (torch.remainder(q[1], P) / h).floor().clamp(0, n - 1).long()) * n + (q[2] * (n - 1)).floor().clamp(0, n - 2).long()
    o = torch.argsort(key)
    assert torch.equal(ip(*[a[o] for a in q]), r[o])          # tile path == direct path
    assert torch.equal(ip(*q, presort=True), r)
    m = 1 << 13
    qn = [a[:m].cpu().numpy() for a in q]
    xn, zn, fn = x.cpu().numpy(), z.cpu().numpy(), f.cpu().numpy()
    want = O.Interpolator3D(xn, xn, zn, fn, "cubic", extrap, period)(*qn)
    up = lambda v: v.astype(np.float64)
    truth = O.Interpolator3D(up(xn), up(xn), up(zn), up(fn), "cubic", extrap, period)(*[up(v) for v in qn])
    assert_parity(r[:m].cpu().numpy(), want, "f32 3-D", truth)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_long_knot_vectors_shrink_the_cta(dtype):
    """5000 x 24 x 40 nodes: the staged knots take most of the shared memory, so the tile
    kernel runs with fewer warps per CTA (sparse stream: long runs, one query per lane) -- and
    a denser, cell-sorted stream on the same table takes the two-per-lane shape.  Parity with
    the oracle and bitwise agreement between the two orders."""
    rng = np.random.default_rng(77)
    nx, ny, nz = 5000, 24, 40
    x = knots_nonuniform(rng, nx, 0.0, 50.0, dtype)
    y = knots_nonuniform(rng, ny, -1.0, 1.0, dtype)
    z = knots_nonuniform(rng, nz, 0.0, 3.0, dtype)
    f = rng.normal(size=(nx, ny, nz)).astype(dtype)
    ip = ipx.Interpolator3D(x, y, z, f, "cubic", extrap=True)
    ref = O.Interpolator3D(x, y, z, f, "cubic", extrap=True)
    for nq in (200_000, 20_000_000 if dtype == np.float64 else 2_000_000):
        xq = rng.uniform(-0.1, 50.1, nq).astype(dtype)
        yq = rng.uniform(-1.02, 1.02, nq).astype(dtype)
        zq = rng.uniform(-0.05, 3.05, nq).astype(dtype)
        key = ((np.searchsorted(x, xq) * ny + np.searchsorted(y, yq)) * nz + np.searchsorted(z, zq))
        o = np.argsort(key, kind="stable")
        r = ip(xq, yq, zq)
        rs = ip(xq[o], yq[o], zq[o])
        assert np.array_equal(rs, r[o])
        m = 4096
        assert_parity(rs[:m], ref(xq[o][:m], yq[o][:m], zq[o][:m]),
                      truth=None if dtype == np.float64 else
                      O.Interpolator3D(upcast(x), upcast(y), upcast(z), upcast(f), "cubic", extrap=True)(
                          upcast(xq[o][:m]), upcast(yq[o][:m]), upcast(zq[o][:m])))


# ------------------------------------------------------------------ tile kernel variants
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("n", [9, 34])
def test_many_queries_every_periodic_mask(ndim, n, dtype):
    """Enough queries for the shared-memory tile kernel in BOTH forms (the functional form packs
    its tables from 2^16 queries up; on periodic axes it hands the kernel PADDED tables, the
    class un-padded ones), for every set of periodic axes -- none / all / all but the last are
    compile-time masks in the kernel, the rest read the flags per query -- in cell-sorted and
    random order (cooperative run copies vs direct gathers), dense and sparse streams (two
    queries or one per lane), with a derivative on one axis.  Checked against the oracle."""
    rng = np.random.default_rng(7 * ndim + n)
    nq = 70_000
    ns = [n + ax for ax in range(ndim)]
    knots = [knots_nonuniform(rng, m, 0.0, 1.0 + ax, dtype) for ax, m in enumerate(ns)]
    f = rng.normal(size=ns).astype(dtype)
    q = [rng.uniform(-0.6 * (1 + ax), 1.7 * (1 + ax), nq).astype(dtype) for ax in range(ndim)]
    key = np.zeros(nq, dtype=np.int64)
    for ax in range(ndim):
        key = key * (ns[ax] + 2) + np.searchsorted(knots[ax], np.mod(q[ax], (1.0 + ax) * 1.05))
    order = np.argsort(key, kind="stable")
    gfun, ofun = (ipx.interp2d, O.interp2d) if ndim == 2 else (ipx.interp3d, O.interp3d)
    gcls, ocls = (ipx.Interpolator2D, O.Interpolator2D) if ndim == 2 else (ipx.Interpolator3D, O.Interpolator3D)
    for mask in range(1 << ndim):
        period = tuple((1.0 + ax) * 1.05 if (mask >> ax) & 1 else None for ax in range(ndim))
        if mask == 0:
            period = None
        extrap = tuple((True, 0.5) for _ in range(ndim))
        deriv = tuple(1 if ax == mask % ndim else 0 for ax in range(ndim))
        for qs in ([a[order] for a in q], q):
            m = 6000  # oracle subsample, spread over the stream
            sel = np.linspace(0, nq - 1, m).astype(np.int64)
            got = gfun(*qs, *knots, f, "cubic", deriv, extrap, period)
            want = ofun(*[a[sel] for a in qs], *knots, f, "cubic", deriv, extrap, period)
            truth = None
            if dtype == np.float32:
                truth = ofun(*[upcast(a[sel]) for a in qs], *[upcast(k) for k in knots], upcast(f), "cubic",
                             deriv, extrap, period)
            assert_parity(got[sel], want, f"functional nd={ndim} n={n} mask={mask}", truth)
            got_c = gcls(*knots, f, "cubic", extrap, period)(*qs, *deriv)
            want_c = ocls(*knots, f, "cubic", extrap, period)(*[a[sel] for a in qs], *deriv)
            truth_c = None
            if dtype == np.float32:
                truth_c = ocls(*[upcast(k) for k in knots], upcast(f), "cubic", extrap, period)(
                    *[upcast(a[sel]) for a in qs], *deriv)
            assert_parity(got_c[sel], want_c, f"class nd={ndim} n={n} mask={mask}", truth_c)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_unaligned_out_and_odd_counts(dtype):
    """The tile kernel stores a lane's two results as one 16-byte (f64) / 8-byte (f32) word when
    `out` allows it: an `out` that starts one element into a buffer, and odd query counts, must
    give the same bits through the element-wise store."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(77)
    n = 20
    k = torch.linspace(0, 1, n, dtype=dtype, device=dev)
    f = torch.randn((n, n, n), dtype=dtype, device=dev, generator=g)
    ip = ipx.Interpolator3D(k, k, k, f, "cubic", ((True, True), (0.5, True), (False, False)), (None, None, 1.25))
    for nq in (1, 2, 63, 64, 65, 4097, 100_001):
        q = [(torch.rand(nq, dtype=dtype, device=dev, generator=g) * 1.4 - 0.2) for _ in range(3)]
        want = ip(*q)
        buf = torch.full((nq + 3,), -5.0, dtype=dtype, device=dev)
        for off in (0, 1, 2):
            buf.fill_(-5.0)
            ret = ip(*q, out=buf[off:off + nq])
            assert torch.equal(torch.nan_to_num(ret, nan=-7.0), torch.nan_to_num(want, nan=-7.0)), (nq, off)
            # nothing written outside the window
            assert torch.all(buf[:off] == -5.0) and torch.all(buf[off + nq:] == -5.0), (nq, off)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_queries_many_periods_out_wrap_to_the_same_bits(ndim, dtype):
    """Periodic axes: a query any number of periods out goes through the out-of-line remainder
    (one division + one fma while the quotient is small, fmod beyond that).  Its result must be
    the Python remainder bit for bit (the reference wraps with `xq % period`, _spline.py:217):
    evaluating q and evaluating np.mod(q, P) -- already inside [0, P), where the kernels do not
    touch the value -- must give identical bits, in the cooperative kernels (many queries) and
    the one-thread-per-query kernel (few)."""
    rng = np.random.default_rng(100 + ndim)
    n = 12
    P = [np.dtype(dtype).type(2 * np.pi), np.dtype(dtype).type(1.3), np.dtype(dtype).type(0.1)][:ndim]
    knots = [np.linspace(0, float(p), n, endpoint=False).astype(dtype) for p in P]
    f = rng.normal(size=(n,) * ndim).astype(dtype)
    cls = (ipx.Interpolator1D, ipx.Interpolator2D, ipx.Interpolator3D)[ndim - 1]
    period = P[0] if ndim == 1 else tuple(P)
    for method in ("cubic", "cubic2", "linear"):
        ip = cls(*knots, f, method, True, period)
        for nq in (300, 70_000):
            q = []
            for p in P:
                span = rng.choice([1.5, 3.0, 40.0, 3e4 if dtype == np.float64 else 900.0], nq)
                v = rng.uniform(-1, 1, nq) * span * float(p)
                k = rng.integers(-60, 60, nq)
                edge = (k * p.astype(np.float64)).astype(dtype)
                v = np.where(rng.random(nq) < 0.05, edge, v.astype(dtype))
                v = np.where(rng.random(nq) < 0.02, np.nextafter(edge, dtype(np.inf)), v)
                v = np.where(rng.random(nq) < 0.02, np.nextafter(edge, dtype(-np.inf)), v)
                v[:4] = [1e15, -1e15, 3e30, -7e29]      # quotients beyond the fast path
                q.append(v.astype(dtype))
            r = [np.mod(v, p) for v, p in zip(q, P)]
            assert all(a.dtype == dtype for a in r)
            # a remainder that rounds up to P itself would wrap once more on the second pass
            keep = np.ones(nq, dtype=bool)
            for a, p in zip(r, P):
                keep &= a < p
            got = np.asarray(ip(*q))
            want = np.asarray(ip(*r))
            assert keep.sum() > 0.9 * nq
            assert np.array_equal(got[keep], want[keep]), (ndim, method, nq,
                                                           np.flatnonzero(got[keep] != want[keep])[:5])

"""Pins for the CPU oracle (oracle/interpax_np.py).  No GPU needed.

The reference cannot be imported here (no jax), so the oracle is pinned against
  * the golden vectors the reference's own tests hold (tests/golden/reference_kat.json,
    transcribed by tests/golden/make_reference_kat.py with file:line citations),
  * the reference's coefficient matrices (hash fixture; element-wise when /root/reference
    is mounted),
  * SciPy's CubicSpline / PchipInterpolator / Akima1DInterpolator / CubicHermiteSpline
    (the reference states its stencils replicate SciPy's: _fd_derivs.py:24-25),
  * np.searchsorted for interval indices,
  * every analytic-function case of the reference's tests/test_interpolate.py with the
    reference's own tolerances.
"""

import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest
import scipy.interpolate as si

from oracle import interpax_np as O

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_kat.json")) as fh:
    KAT = json.load(fh)


def hermite(x, y, method, xi, nu=0, **kw):
    """What the reference's scipy-like classes compute: Hermite cubic from approx_df."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    fx = O.approx_df(x, y, method, 0, **kw)
    return O.interp1d(np.asarray(xi, dtype=float), x, y, "cubic", nu, True, None, fx=fx)


# ---------------------------------------------------------------- coefficient matrices
@pytest.mark.parametrize("name", ["A_CUBIC", "A_BICUBIC", "A_TRICUBIC"])
def test_coef_matrices_match_reference_hash(name):
    a = np.ascontiguousarray(getattr(O, name).astype(np.int64))
    g = KAT["coefs"][name]
    assert list(a.shape) == g["shape"]
    assert hashlib.sha256(a.tobytes()).hexdigest() == g["sha256"]
    assert int(np.count_nonzero(a)) == g["nnz"]
    assert [int(v) for v in a.sum(axis=1)] == g["rowsum"]


@pytest.mark.skipif(
    not os.path.exists("/root/reference/interpax/_coefs.py"), reason="reference not mounted"
)
def test_coef_matrices_match_reference_file():
    spec = importlib.util.spec_from_file_location(
        "_ref_coefs", "/root/reference/interpax/_coefs.py"
    )
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in ("A_CUBIC", "A_BICUBIC", "A_TRICUBIC"):
        assert np.array_equal(getattr(mod, name), getattr(O, name))


# ---------------------------------------------------------------- reference golden vectors
def test_akima_golden():
    g = KAT["akima_eval"]
    np.testing.assert_allclose(hermite(g["x"], g["y"], "akima", g["xi"]), g["yi"], rtol=g["rtol"])
    # 2-D / 3-D variants (test_scipy.py:93-155): extra trailing columns scale linearly
    y = np.asarray(g["y"])
    Y = np.stack([y, 2 * y, 3 * y, 4 * y], axis=1).reshape(11, 2, 2)
    fx = O.approx_df(np.asarray(g["x"]), Y, "akima", 0)
    out = O.interp1d(np.asarray(g["xi"]), np.asarray(g["x"]), Y, "cubic", 0, True, fx=fx)
    yi = np.asarray(g["yi"])
    np.testing.assert_allclose(
        out, np.stack([yi, 2 * yi, 3 * yi, 4 * yi], axis=1).reshape(13, 2, 2), rtol=g["rtol"]
    )


def test_akima_degenerate():
    g = KAT["akima_degenerate"]
    x = np.asarray(g["x"])
    y = np.vstack((x, x**2)).T
    xe = np.asarray(g["x_eval"])
    fx = O.approx_df(x, y, "akima", 0)
    out = O.interp1d(xe, x, y, "cubic", 0, True, fx=fx)
    np.testing.assert_allclose(out, np.vstack((xe, xe**2)).T)


def test_pchip_nag():
    g = KAT["pchip_nag"]
    np.testing.assert_allclose(
        hermite(g["x"], g["y"], "monotonic", g["xi"]), g["yi"], rtol=0.0, atol=g["atol"]
    )


def test_pchip_cast_endslopes_two_three_points():
    g = KAT["pchip_cast"]
    xx = np.arange(100)
    a = hermite(g["x"], g["y"], "monotonic", xx)
    b = hermite(np.asarray(g["x"]) * 1.0, np.asarray(g["y"]) * 1.0, "monotonic", xx)
    np.testing.assert_allclose(a, b, atol=1e-14, rtol=1e-14)
    g = KAT["pchip_endslopes"]
    for y in (g["y1"], g["y2"]):
        d = O.approx_df(np.asarray(g["x"]), np.asarray(y), "monotonic", 0)
        assert d[0] != 0 and d[-1] != 0
    g = KAT["pchip_two_points"]
    xs = np.linspace(0, 1, 11)
    np.testing.assert_allclose(hermite(g["x"], g["y"], "monotonic", xs), 2 * xs, atol=1e-15)
    g = KAT["pchip_three_points"]
    np.testing.assert_allclose(hermite(g["x"], g["y"], "monotonic", g["xi"], 0), g["d0"], atol=1e-6)
    np.testing.assert_allclose(hermite(g["x"], g["y"], "monotonic", g["xi"], 1), g["d1"], atol=1e-6)
    # all zeros (test_scipy.py:698-708)
    assert np.array_equal(hermite(np.arange(10), np.zeros(10), "monotonic", np.linspace(0, 9, 101)), np.zeros(101))


def test_monotonic_heaviside_signs():
    # tests/test_interpolate.py:136-149
    x = np.linspace(-4, 5, 10)
    f = np.heaviside(x - 1.5, 0) + 0.1 * x
    xq = np.linspace(-4, 5, 1000)
    dfc = O.interp1d(xq, x, f, derivative=1, method="cubic")
    dfm = O.interp1d(xq, x, f, derivative=1, method="monotonic")
    dfm0 = O.interp1d(xq, x, f, derivative=1, method="monotonic-0")
    assert dfc.min() < 0 and dfm.min() > 0 and dfm0.min() >= 0
    np.testing.assert_allclose(dfm0[np.array([0, -1])], 0, atol=1e-12)


# ---------------------------------------------------------------- SciPy pins
RNG = np.random.default_rng(7)
XS = np.sort(RNG.uniform(0, 5, 40))
YS = RNG.normal(size=(40, 3))


def test_scipy_pchip_slopes_and_values():
    p = si.PchipInterpolator(XS, YS, axis=0)
    np.testing.assert_allclose(O.approx_df(XS, YS, "monotonic", 0), p(XS, 1), rtol=1e-12, atol=1e-12)
    xi = np.linspace(XS[0], XS[-1], 501)
    fx = O.approx_df(XS, YS, "monotonic", 0)
    for nu in (0, 1, 2, 3):
        sc = np.abs(p(xi, nu)).max()
        np.testing.assert_allclose(
            O.interp1d(xi[:-1], XS, YS, "cubic", nu, True, fx=fx), p(xi[:-1], nu), rtol=1e-9, atol=1e-11 * sc
        )


def test_scipy_akima_slopes():
    p = si.Akima1DInterpolator(XS, YS, axis=0)
    # per-column maxima differ from the reference's whole-array threshold only in
    # degenerate cases, not on random data
    np.testing.assert_allclose(O.approx_df(XS, YS, "akima", 0), p(XS, 1), rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize(
    "bc",
    ["not-a-knot", "natural", "clamped", ((1, 2.0), (2, -1.0)), ((2, -1.0), "not-a-knot"), ("natural", (1, 2.0))],
)
@pytest.mark.parametrize("n", [2, 3, 4, 40])
def test_scipy_cubicspline_slopes(bc, n):
    x, y = XS[:n], YS[:n, 0]
    if n == 2 and not isinstance(bc, str) and "not-a-knot" in bc:
        pytest.skip("scipy rejects this mix at n=2")
    if isinstance(bc, str):
        bco = bc
    else:
        bco = tuple(b if isinstance(b, str) else (b[0], np.asarray(b[1])) for b in bc)
    ref = si.CubicSpline(x, y, bc_type=bc)(x, 1)
    np.testing.assert_allclose(O.approx_df(x, y, "cubic2", 0, bc_type=bco), ref, rtol=1e-10, atol=1e-11)


def test_scipy_cubicspline_batched_axis():
    Y = RNG.normal(size=(2, 12, 3))
    x = np.sort(RNG.uniform(0, 1, 12))
    for axis in (1,):
        ref = si.CubicSpline(x, Y, axis=axis)(x, 1)
        np.testing.assert_allclose(O.approx_df(x, Y, "cubic2", axis), ref, rtol=1e-10, atol=1e-11)


def test_scipy_hermite_eval_all_orders():
    fx = RNG.normal(size=YS.shape)
    p = si.CubicHermiteSpline(XS, YS, fx, axis=0)
    xi = RNG.uniform(XS[0], XS[-1], 1000)
    for nu in range(4):
        ref = p(xi, nu)
        np.testing.assert_allclose(
            O.interp1d(xi, XS, YS, "cubic", nu, True, fx=fx), ref, rtol=1e-9, atol=1e-11 * np.abs(ref).max()
        )
    assert np.array_equal(O.interp1d(xi, XS, YS, "cubic", 4, True, fx=fx), np.zeros((1000, 3)))
    # knots are reproduced exactly (test_scipy.py:898-904)
    np.testing.assert_allclose(O.interp1d(XS, XS, YS, "cubic", 0, True, fx=fx), YS, rtol=1e-15)


def test_scipy_regular_grid_linear_nearest():
    x = np.sort(RNG.uniform(0, 1, 9)); y = np.sort(RNG.uniform(0, 2, 7)); z = np.sort(RNG.uniform(-1, 1, 8))
    f = RNG.normal(size=(9, 7, 8))
    q = np.stack([RNG.uniform(x[0], x[-1], 500), RNG.uniform(y[0], y[-1], 500), RNG.uniform(z[0], z[-1], 500)], 1)
    ref = si.RegularGridInterpolator((x, y, z), f, method="linear")(q)
    np.testing.assert_allclose(O.interp3d(q[:, 0], q[:, 1], q[:, 2], x, y, z, f, "linear"), ref, rtol=1e-10, atol=1e-12)
    ref2 = si.RegularGridInterpolator((x, y), f[:, :, 0], method="linear")(q[:, :2])
    np.testing.assert_allclose(O.interp2d(q[:, 0], q[:, 1], x, y, f[:, :, 0], "linear"), ref2, rtol=1e-10, atol=1e-12)


def test_bin_index_is_searchsorted_right():
    x = np.array([0.0, 0.5, 0.5, 1.0, 2.0, 2.0, 3.0])
    q = np.array([-1.0, 0.0, -0.0, 0.25, 0.5, 0.75, 1.0, 2.0, 2.5, 3.0, 4.0, np.nan, np.inf, -np.inf])
    want = np.array([1, 1, 1, 1, 3, 3, 4, 6, 6, 6, 6, 6, 6, 1])
    assert np.array_equal(O.bin_index(x, q), want)


# ---------------------------------------------------------------- the reference's analytic tests
TOL_1D = {
    "nearest": (1e-2, 1e-1), "linear": (1e-4, 1e-3), "cubic": (1e-6, 1e-5), "cubic2": (1e-6, 1e-5),
    "cardinal": (1e-6, 1e-5), "catmull-rom": (1e-6, 1e-5), "monotonic": (1e-4, 1e-3),
    "monotonic-0": (1e-4, 1.5e-2), "akima": (1e-6, 2e-5),
}


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("scalar", [False, True])
def test_reference_interp1d_cases(dtype, scalar):
    # tests/test_interpolate.py:26-82
    xp = np.real(np.linspace(0, 2 * np.pi, 100).astype(dtype))
    cplx = np.iscomplexobj(np.zeros(1, dtype))
    f = (lambda x: np.sin(x) + 1j * np.sin(x)) if cplx else (lambda x: np.sin(x))
    fp = f(xp).astype(dtype)
    x = 0.0 if scalar else np.real(np.linspace(0, 2 * np.pi, 10000).astype(dtype))
    for interp in (O.interp1d, lambda xq, *a, **k: O.Interpolator1D(*a, **k)(xq)):
        for m, (rtol, atol) in TOL_1D.items():
            fq = interp(x, xp, fp, method=m)
            assert fq.dtype == np.result_type(x, xp, fp), m
            np.testing.assert_allclose(fq, f(x), rtol=rtol, atol=atol, err_msg=m)


def test_reference_interp1d_extrap_periodic():
    # tests/test_interpolate.py:117-133
    xp = np.linspace(0, 2 * np.pi, 200)
    x = np.linspace(-1, 2 * np.pi + 1, 10000)
    fp = np.sin(xp)
    fq = O.interp1d(x, xp, fp, method="cubic", extrap=False)
    assert np.isnan(fq[0]) and np.isnan(fq[-1])
    fq = O.interp1d(x, xp, fp, method="cubic", extrap=True)
    assert not np.isnan(fq[0]) and not np.isnan(fq[-1])
    fq = O.interp1d(x, xp, fp, method="cubic", period=2 * np.pi)
    np.testing.assert_allclose(fq, np.sin(x), rtol=1e-6, atol=1e-2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("scalar", [False, True])
def test_reference_interp2d_cases(dtype, scalar):
    # tests/test_interpolate.py:155-235
    xp = np.linspace(0, 3 * np.pi, 99).astype(dtype)
    yp = np.linspace(0, 2 * np.pi, 40).astype(dtype)
    xxp, yyp = np.meshgrid(xp, yp, indexing="ij")
    f = lambda x, y: np.sin(x) * np.cos(y)
    fp = f(xxp, yyp).astype(dtype)
    if scalar:
        x, y = 0.0, 0.0
    else:
        x = np.linspace(0, 3 * np.pi, 1000).astype(dtype)
        y = np.linspace(0, 2 * np.pi, 1000).astype(dtype)
    per = (2 * np.pi, 2 * np.pi)
    tol = {"nearest": (1e-2, 1), "linear": (1e-4, 1e-2), "cubic": (1e-5, 2e-3), "cubic2": (1e-5, 2e-3),
           "catmull-rom": (1e-5, 2e-3), "cardinal": (1e-5, 2e-3), "akima": (1e-5, 2e-3),
           "monotonic": (1e-2, 1e-3), "monotonic-0": (1e-2, 2e-2)}
    for interp in (O.interp2d, lambda xq, yq, *a, **k: O.Interpolator2D(*a, **k)(xq, yq)):
        for m, (rtol, atol) in tol.items():
            fq = interp(x, y, xp, yp, fp, method=m, period=per)
            assert fq.dtype == np.result_type(x, y, xp, yp, fp)
            np.testing.assert_allclose(fq, f(x, y), rtol=rtol, atol=atol, err_msg=m)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reference_interp3d_cases(dtype):
    # tests/test_interpolate.py:262-337
    xp = np.linspace(0, np.pi, 20).astype(dtype)
    yp = np.linspace(0, 2 * np.pi, 30).astype(dtype)
    zp = np.linspace(0, 3, 25).astype(dtype)
    xxp, yyp, zzp = np.meshgrid(xp, yp, zp, indexing="ij")
    f = lambda x, y, z: np.sin(x) * np.cos(y) * z**2
    fp = f(xxp, yyp, zzp).astype(dtype)
    x = np.linspace(0, np.pi, 1000).astype(dtype)
    y = np.linspace(0, 2 * np.pi, 1000).astype(dtype)
    z = np.linspace(0, 3, 1000).astype(dtype)
    tol = {"nearest": (1e-2, 1), "linear": (1e-3, 1e-1), "cubic": (1e-5, 5.5e-3), "cubic2": (1e-5, 5.5e-3),
           "catmull-rom": (1e-5, 5.5e-3), "cardinal": (1e-5, 5.5e-3), "monotonic": (1e-2, 1e-2),
           "monotonic-0": (1e-2, 0.3)}
    for interp in (O.interp3d, lambda a, b, c, *r, **k: O.Interpolator3D(*r, **k)(a, b, c)):
        for m, (rtol, atol) in tol.items():
            fq = interp(x, y, z, xp, yp, zp, fp, method=m)
            assert fq.dtype == np.result_type(x, y, z, xp, yp, zp, fp)
            np.testing.assert_allclose(fq, f(x, y, z), rtol=rtol, atol=atol, err_msg=m)


def test_reference_vector_valued_and_extrap_float():
    # tests/test_interpolate.py:85-114 (1-D trailing batch) and :640-649 (float fill)
    xp = np.linspace(0, 2 * np.pi, 100)
    x = np.linspace(0, 2 * np.pi, 300)[10:-10]
    f = lambda x: np.array([np.sin(x), np.cos(x)])
    fp = f(xp).T
    for m, (rtol, atol) in TOL_1D.items():
        if m == "akima":
            continue
        np.testing.assert_allclose(O.interp1d(x, xp, fp, method=m), f(x).T, rtol=rtol, atol=max(atol, 1e-2 if m == "monotonic-0" else 0))
    x = np.linspace(0, 10, 10); y = np.linspace(0, 8, 8); z = np.zeros((10, 8)) + 1.0
    ip = O.Interpolator2D(x, y, z, extrap=0.0)
    np.testing.assert_allclose(ip(4.5, 5.3), 1.0)
    np.testing.assert_allclose(ip(-4.5, 5.3), 0.0)
    np.testing.assert_allclose(ip(4.5, -5.3), 0.0)


def test_separable_form_equals_dense_form():
    """the dense 64x64 contraction is a tensor product of 4x4 ones; the kernels
    evaluate the separable form, the oracle the dense one."""
    rng = np.random.default_rng(0)
    F = rng.normal(size=64)
    t = rng.uniform(size=3)
    coef = (O.A_TRICUBIC @ F).reshape((4, 4, 4), order="F")
    dense = np.einsum("ijk,i,j,k->", coef, *[np.array([1, s, s * s, s**3]) for s in t])
    w = [np.array([1, s, s * s, s**3]) @ O.A_CUBIC for s in t]
    sep = 0.0
    for ff, name in enumerate(O.TABLES_3D):
        a, b, c = (1 if s in name[1:] else 0 for s in "xyz")
        for kk in (0, 1):
            for jj in (0, 1):
                for ii in (0, 1):
                    sep += F[8 * ff + 4 * kk + 2 * jj + ii] * w[0][2 * a + ii] * w[1][2 * b + jj] * w[2][2 * c + kk]
    assert abs(dense - sep) < 1e-12

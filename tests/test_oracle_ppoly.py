"""Pins of oracle/ppoly_np.py (the NumPy restatement of interpax/_ppoly.py) against SciPy's
classes, which the reference's own tests (tests/test_scipy.py) use as ground truth, plus the
closed-form cases of that file restated."""

import numpy as np
import pytest
import scipy.interpolate as si

from oracle import ppoly_np as P


def rnd(seed=0):
    return np.random.default_rng(seed)


@pytest.mark.parametrize("k", [1, 2, 4, 6])
@pytest.mark.parametrize("tail", [(), (3,), (2, 3)])
def test_ppoly_call_matches_scipy(k, tail):
    g = rnd(k)
    m = 9
    x = np.sort(g.uniform(0, 4, m + 1))
    c = g.normal(size=(k, m) + tail)
    xq = g.uniform(-0.5, 4.5, 200)
    for extrapolate in (True, False, "periodic"):
        ours, ref = P.PPoly(c, x, extrapolate), si.PPoly(c, x, extrapolate)
        for nu in range(0, k + 2):
            np.testing.assert_allclose(ours(xq, nu), ref(xq, nu), rtol=1e-12, atol=1e-12)


def test_ppoly_interval_convention_and_shapes():
    # test_scipy.py:287-293 (test_simple), :199-215 (test_shape), half-open intervals
    c = np.array([[1, 4], [2, 5], [3, 6]])
    x = np.array([0, 0.5, 1])
    p = P.PPoly(c, x)
    np.testing.assert_allclose(p(0.3), 1 * 0.3**2 + 2 * 0.3 + 3)
    np.testing.assert_allclose(p(0.7), 4 * (0.7 - 0.5) ** 2 + 5 * (0.7 - 0.5) + 6)
    np.testing.assert_allclose(p(0.5), 6.0)          # the breakpoint belongs to the right piece
    np.testing.assert_allclose(p(1.0), 4 * 0.25 + 5 * 0.5 + 6)  # last interval is closed
    g = rnd(1)
    c = g.normal(size=(8, 12, 5, 6, 7))
    x = np.sort(g.uniform(size=13))
    xp = g.uniform(size=(3, 4))
    assert P.PPoly(c, x)(xp).shape == (3, 4, 5, 6, 7)
    assert np.isnan(P.PPoly(c, x, extrapolate=False)(np.array([-1.0, 2.0]))).all()


def test_ppoly_axis_matches_scipy():
    # test_scipy.py:229-260
    g = rnd(2)
    c = g.normal(size=(3, 4, 5, 6, 7, 8))
    xq = g.uniform(size=(1, 2))
    for axis in (0, 1, 2, 3):
        m = c.shape[axis + 1]
        x = np.sort(g.uniform(size=m + 1))
        ours, ref = P.PPoly(c, x, axis=axis), si.PPoly(c, x, axis=axis)
        assert ours.c.shape == ref.c.shape
        np.testing.assert_allclose(ours(xq), ref(xq), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(ours(xq, 1), ref(xq, 1), rtol=1e-12, atol=1e-13)


def test_ppoly_complex_coefficients():
    # test_scipy.py:217-227
    x = np.sort(rnd(3).uniform(size=11))
    c = rnd(4).normal(size=(4, 10)) * (1 + 2j)
    xq = rnd(5).uniform(size=50)
    ours, ref = P.PPoly(c, x), si.PPoly(c, x)
    np.testing.assert_allclose(ours(xq), ref(xq), rtol=1e-12)
    np.testing.assert_allclose(ours(xq, 1), ref(xq, 1), rtol=1e-12)


@pytest.mark.parametrize("k", [1, 3, 5])
def test_derivative_antiderivative_integrate_match_scipy(k):
    g = rnd(10 + k)
    m = 7
    x = np.sort(g.uniform(0, 3, m + 1))
    c = g.normal(size=(k, m, 2))
    ours, ref = P.PPoly(c, x), si.PPoly(c, x)
    xq = g.uniform(-0.2, 3.2, 100)
    for nu in (1, 2, 3):
        np.testing.assert_allclose(ours.derivative(nu).c, ref.derivative(nu).c, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(ours.antiderivative(nu).c, ref.antiderivative(nu).c, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(ours.antiderivative(nu)(xq), ref.antiderivative(nu)(xq), rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(ours.derivative(-nu).c, ref.antiderivative(nu).c, rtol=1e-12, atol=1e-12)
    # continuity of the antiderivative at the breakpoints (test_scipy.py:472-484)
    ia = ours.antiderivative()
    eps = 1e-9 * (x[1:-1] + 1)
    np.testing.assert_allclose(ia(x[1:-1] - eps), ia(x[1:-1] + eps), rtol=1e-6, atol=1e-7)
    for a, b in ((0.3, 2.1), (2.5, 0.1), (-0.4, 3.3)):
        np.testing.assert_allclose(ours.integrate(a, b), ref.integrate(a, b), rtol=1e-11, atol=1e-12)
    per_o, per_r = P.PPoly(c, x, "periodic"), si.PPoly(c, x, "periodic")
    for a, b in ((0.3, 2.1), (-1.0, 9.5), (7.0, -3.0)):
        np.testing.assert_allclose(per_o.integrate(a, b), per_r.integrate(a, b), rtol=1e-10, atol=1e-11)


def test_hermite_family_matches_scipy():
    g = rnd(20)
    n = 15
    x = np.sort(g.uniform(0, 5, n))
    y = g.normal(size=(n, 3))
    d = g.normal(size=(n, 3))
    xq = g.uniform(-0.3, 5.3, 300)
    pairs = [
        (P.CubicHermiteSpline(x, y, d), si.CubicHermiteSpline(x, y, d)),
        (P.PchipInterpolator(x, y), si.PchipInterpolator(x, y)),
        (P.Akima1DInterpolator(x, y), si.Akima1DInterpolator(x, y, extrapolate=True)),
        (P.CubicSpline(x, y), si.CubicSpline(x, y)),
        (P.CubicSpline(x, y, bc_type="natural"), si.CubicSpline(x, y, bc_type="natural")),
        (P.CubicSpline(x, y, bc_type="clamped"), si.CubicSpline(x, y, bc_type="clamped")),
        (P.CubicSpline(x, y, bc_type=((1, d[0]), (2, d[-1]))), si.CubicSpline(x, y, bc_type=((1, d[0]), (2, d[-1])))),
    ]
    for ours, ref in pairs:
        np.testing.assert_allclose(ours.c, ref.c, rtol=1e-9, atol=1e-9)
        for nu in (0, 1, 2):
            np.testing.assert_allclose(ours(xq, nu), ref(xq, nu), rtol=1e-9, atol=1e-9)
    # axis != 0
    y2 = np.moveaxis(y, 0, 1)
    np.testing.assert_allclose(P.PchipInterpolator(x, y2, axis=1)(xq), si.PchipInterpolator(x, y2, axis=1)(xq),
                               rtol=1e-10, atol=1e-10)


def test_pchip_known_answers():
    # test_scipy.py:653-686 (NAG example) and :710-717 (two points)
    dataStr = """
          7.99   0.00000E+0
          8.09   0.27643E-4
          8.19   0.43750E-1
          8.70   0.16918E+0
          9.20   0.46943E+0
         10.00   0.94374E+0
         12.00   0.99864E+0
         15.00   0.99992E+0
         20.00   0.99999E+0
    """
    data = np.loadtxt(dataStr.strip().splitlines())
    pch = P.PchipInterpolator(data[:, 0], data[:, 1])
    resultStr = """
       7.9900       0.0000
       9.1910       0.4640
      10.3920       0.9645
      11.5930       0.9965
      12.7940       0.9992
      13.9950       0.9998
      15.1960       0.9999
      16.3970       1.0000
      17.5980       1.0000
      18.7990       1.0000
      20.0000       1.0000
    """
    result = np.loadtxt(resultStr.strip().splitlines())
    np.testing.assert_allclose(result[:, 1], pch(result[:, 0]), rtol=0.0, atol=5e-5)
    p = P.PchipInterpolator([0, 1], [0, 2])
    np.testing.assert_allclose(p(np.linspace(0, 1, 11)), 2 * np.linspace(0, 1, 11), atol=1e-15)


def test_constructor_errors():
    # test_scipy.py:176-184
    with pytest.raises(ValueError):
        P.PPoly(np.array([[1, 2], [3, 4]], dtype=float), np.array([0, 1, 0.5]))
    with pytest.raises(ValueError):
        P.PPoly(np.array([1, 2]), np.array([0, 1]))
    with pytest.raises(ValueError):
        P.CubicSpline([0, 1, 1], [1, 2, 3])
    with pytest.raises(ValueError):
        P.CubicHermiteSpline([0, 1, 2], [1, 2, 3], [1, 2])

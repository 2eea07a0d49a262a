"""GPU parity of the piecewise-polynomial family (interpax_b200/ppoly.py -> C ABI ->
csrc/ipx_ppoly.cu) against the NumPy restatement oracle/ppoly_np.py, which
tests/test_oracle_ppoly.py pins against SciPy.  float64: rtol 1e-12 (paired with
atol = rtol * max|oracle|), float32: 1e-5; interval convention bit-exact (checked through
piecewise-constant polynomials whose value IS the piece index)."""

import numpy as np
import pytest
import torch

import interpax_b200 as ipx
from oracle import ppoly_np as O

pytestmark = pytest.mark.gpu


def close(got, want, rtol=1e-12):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    fin = np.isfinite(want)
    if fin.any():
        scale = np.abs(want[fin]).max()
        np.testing.assert_allclose(got[fin], want[fin], rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("k", [1, 2, 4, 7])
@pytest.mark.parametrize("tail", [(), (3,), (2, 5)])
def test_ppoly_call_parity(k, tail, dtype):
    g = np.random.default_rng(100 + k)
    m = 23
    x = np.sort(g.uniform(0, 4, m + 1)).astype(dtype)
    c = g.normal(size=(k, m) + tail).astype(dtype)
    xq = np.concatenate([g.uniform(-0.5, 4.5, 777), x, [np.nan, np.inf, -np.inf]]).astype(dtype)
    rtol = 1e-12 if dtype == np.float64 else 1e-5
    for extrapolate in (True, False, "periodic"):
        ours, ref = ipx.PPoly(c, x, extrapolate), O.PPoly(c, x, extrapolate)
        for nu in range(0, k + 2):
            with np.errstate(invalid="ignore"):
                want = ref(xq, nu)
            close(ours(xq, nu), want, rtol)
    # torch in -> torch out on the device, any query shape, scalar queries
    p = ipx.PPoly(torch.as_tensor(c, device="cuda"), torch.as_tensor(x, device="cuda"))
    r = p(torch.as_tensor(xq[:20].reshape(4, 5), device="cuda"))
    assert r.is_cuda and tuple(r.shape) == (4, 5) + tail
    close(r.cpu().numpy(), O.PPoly(c, x)(xq[:20].reshape(4, 5)), rtol)
    assert ipx.PPoly(c, x)(float(x[3])).shape == tail


def test_ppoly_interval_convention_is_exact():
    # piecewise-constant pieces: the value IS the piece index; breakpoints belong to the right
    # piece, the last interval is closed, duplicates in x resolve like searchsorted(side="right")
    g = np.random.default_rng(7)
    x = np.sort(g.uniform(0, 1, 2000))
    x[500] = x[501]
    x[900:903] = x[900]
    m = x.size - 1
    c = np.arange(m, dtype=np.float64)[None, :]
    xq = np.concatenate([g.uniform(-0.1, 1.1, 100000), x, np.nextafter(x, 2), np.nextafter(x, -2)])
    got = ipx.PPoly(c, x)(xq)
    want = np.clip(np.searchsorted(x, xq, side="right"), 1, m) - 1
    assert np.array_equal(got, want.astype(np.float64))


def test_ppoly_axis_and_complex():
    g = np.random.default_rng(8)
    c = g.normal(size=(3, 4, 5, 6, 7))
    xq = g.uniform(size=(2, 3))
    for axis in (0, 1, 2):
        x = np.sort(g.uniform(size=c.shape[axis + 1] + 1))
        ours, ref = ipx.PPoly(c, x, axis=axis), O.PPoly(c, x, axis=axis)
        assert ours.c.shape == ref.c.shape and ours.axis == ref.axis
        close(ours(xq), ref(xq))
        close(ours(xq, 1), ref(xq, 1))
    x = np.sort(g.uniform(size=11))
    cc = g.normal(size=(4, 10, 2)) * (1 + 2j)
    ours, ref = ipx.PPoly(cc, x), O.PPoly(cc, x)
    close(ours(xq), ref(xq))
    close(ours(xq, 2), ref(xq, 2))
    close(ours.derivative().c, ref.derivative().c)
    close(ours.antiderivative().c, ref.antiderivative().c)


@pytest.mark.parametrize("k", [1, 3, 5])
def test_derivative_antiderivative_integrate_parity(k):
    g = np.random.default_rng(20 + k)
    m = 31
    x = np.sort(g.uniform(0, 3, m + 1))
    c = g.normal(size=(k, m, 2))
    ours, ref = ipx.PPoly(c, x), O.PPoly(c, x)
    xq = g.uniform(-0.2, 3.2, 500)
    for nu in (0, 1, 2, 3, 6):
        close(ours.derivative(nu).c, ref.derivative(nu).c)
        close(ours.derivative(nu)(xq), ref.derivative(nu)(xq))
    for nu in (1, 2, 3):
        close(ours.antiderivative(nu).c, ref.antiderivative(nu).c)
        close(ours.derivative(-nu).c, ref.antiderivative(nu).c)
        close(ours.antiderivative(nu)(xq), ref.antiderivative(nu)(xq))
        # derivative undoes antiderivative (test_scipy.py:443-470)
        close(ours.antiderivative(nu).derivative(nu)(xq), ref(xq), 1e-10)
    for a, b in ((0.3, 2.1), (2.5, 0.1), (-0.4, 3.3)):
        close(ours.integrate(a, b), ref.integrate(a, b), 1e-11)
    assert np.isnan(ipx.PPoly(c, x, extrapolate=False).integrate(-1.0, 1.0)).all()
    per_o, per_r = ipx.PPoly(c, x, "periodic"), O.PPoly(c, x, "periodic")
    assert per_o.antiderivative().extrapolate is False
    for a, b in ((0.3, 2.1), (-1.0, 9.5), (7.0, -3.0), (2.9, 3.4)):
        close(per_o.integrate(a, b), per_r.integrate(a, b), 1e-10)
    # subclasses survive derivative() (test_scipy.py:263-283)
    cs = ipx.CubicSpline(x, g.normal(size=x.size))
    assert type(cs.derivative()) is ipx.CubicSpline and cs.derivative().c.shape[0] == 3


@pytest.mark.parametrize("tail,axis", [((), 0), ((3,), 0), ((4, 2), 0), ((3,), 1)])
def test_spline_family_parity(tail, axis):
    g = np.random.default_rng(30)
    n = 41
    # knot spacing bounded below: c = a / dx^3 amplifies the rounding of the 4-term sums in
    # A_CUBIC @ F by 1/dx^3, and neither the oracle's nor XLA's summation order is canonical
    x = np.cumsum(g.uniform(0.06, 0.19, n))
    x = (x - x[0]) * (5.0 / (x[-1] - x[0]))
    shape = (n,) + tail if axis == 0 else tail + (n,)
    y = g.normal(size=shape)
    d = g.normal(size=shape)
    xq = g.uniform(-0.3, 5.3, 1000)
    bc0 = g.normal(size=tail)
    cases = [
        (ipx.CubicHermiteSpline(x, y, d, axis=axis), O.CubicHermiteSpline(x, y, d, axis=axis)),
        (ipx.PchipInterpolator(x, y, axis=axis), O.PchipInterpolator(x, y, axis=axis)),
        (ipx.Akima1DInterpolator(x, y, axis=axis), O.Akima1DInterpolator(x, y, axis=axis)),
        (ipx.CubicSpline(x, y, axis=axis), O.CubicSpline(x, y, axis=axis)),
        (ipx.CubicSpline(x, y, axis=axis, bc_type="natural"), O.CubicSpline(x, y, axis=axis, bc_type="natural")),
        (ipx.CubicSpline(x, y, axis=axis, bc_type=((1, bc0), (2, bc0))),
         O.CubicSpline(x, y, axis=axis, bc_type=((1, bc0), (2, bc0)))),
    ]
    for ours, ref in cases:
        close(ours.c, ref.c, 1e-11)
        for nu in (0, 1, 2, 3):
            close(ours(xq, nu), ref(xq, nu), 1e-11)
        close(ours(xq, extrapolate=False), ref(xq, extrapolate=False), 1e-11)
    # float32 data with the reference's float64 knots give float64 coefficients.  The reference
    # (and the oracle) difference the float32 data before promoting; the CUDA path widens the
    # data first, so the two agree to float32 rounding only -- and exactly with the float64 run
    # on the widened data.
    y32 = y.astype(np.float32)
    p = ipx.PchipInterpolator(x, y32, axis=axis)
    assert p.c.dtype == np.float64
    close(p(xq), O.PchipInterpolator(x, y32, axis=axis)(xq), 1e-5)
    close(p(xq), O.PchipInterpolator(x, y32.astype(np.float64), axis=axis)(xq), 1e-11)


def test_spline_known_answers_and_errors():
    # PCHIP NAG example (test_scipy.py:653-686) through the CUDA path
    xs = np.array([7.99, 8.09, 8.19, 8.70, 9.20, 10.00, 12.00, 15.00, 20.00])
    ys = np.array([0.0, 0.27643e-4, 0.43750e-1, 0.16918, 0.46943, 0.94374, 0.99864, 0.99992, 0.99999])
    xr = np.array([7.99, 9.191, 10.392, 11.593, 12.794, 13.995, 15.196, 16.397, 17.598, 18.799, 20.0])
    yr = np.array([0.0, 0.4640, 0.9645, 0.9965, 0.9992, 0.9998, 0.9999, 1.0, 1.0, 1.0, 1.0])
    np.testing.assert_allclose(ipx.PchipInterpolator(xs, ys)(xr), yr, rtol=0, atol=5e-5)
    # a cubic is reproduced by the not-a-knot spline; sin to spline accuracy
    x = np.linspace(0, 2, 30)
    np.testing.assert_allclose(ipx.CubicSpline(x, x**3 - x)(np.linspace(0, 2, 101)),
                               np.linspace(0, 2, 101) ** 3 - np.linspace(0, 2, 101), atol=1e-12)
    with pytest.raises(ValueError):
        ipx.PPoly(np.array([[1.0, 2], [3, 4]]), np.array([0, 1, 0.5]))
    with pytest.raises(ValueError):
        ipx.PPoly(np.array([1.0, 2]), np.array([0.0, 1]))
    with pytest.raises(ValueError):
        ipx.CubicSpline([0, 1, 1], [1, 2, 3])
    with pytest.raises(ValueError):
        ipx.CubicHermiteSpline([0, 1, 2], [1, 2, 3], [1, 2])
    with pytest.raises(ValueError):
        ipx.CubicSpline([0, 1, 2], [1, np.nan, 3])
    with pytest.raises(NotImplementedError):
        ipx.PPoly(np.ones((2, 2)), np.array([0.0, 1, 2])).roots()


def test_ppoly_large_stream():
    # 1e6 breakpoints, 2e7 queries: spot parity and the closed-form value of a spline of a cubic
    n = 1_000_000
    x = torch.linspace(0, 10, n, dtype=torch.float64, device="cuda")
    y = x**3 - 2 * x
    cs = ipx.CubicSpline(x, y)
    g = torch.Generator(device="cuda").manual_seed(5)
    xq = torch.rand(20_000_000, dtype=torch.float64, device="cuda", generator=g) * 10
    r = cs(xq)
    assert float((r - (xq**3 - 2 * xq)).abs().max()) < 1e-9
    sub = xq[:4096].cpu().numpy()
    ref = O.CubicSpline(x.cpu().numpy(), y.cpu().numpy())
    close(r[:4096].cpu().numpy(), ref(sub), 1e-11)
    close(cs(xq[:4096], 1).cpu().numpy(), ref(sub, 1), 1e-10)


def test_ppoly_piece_major_copy_is_bitwise_neutral():
    # many queries on a scalar-valued polynomial run from the (m, k) copy written by
    # ipx_ppoly_pack_*; short calls read the reference's (k, m) layout: same bits either way
    g = np.random.default_rng(11)
    m, k = 257, 5
    x = np.sort(g.uniform(-1, 1, m + 1))
    c = g.normal(size=(k, m))
    xq = g.uniform(-1.2, 1.2, 8 * m)
    p = ipx.PPoly(c, x)
    for nu in (0, 1, 3):
        whole = p(xq, nu)                                  # >= 4 m queries: piece-major
        parts = np.concatenate([p(xq[i:i + m], nu) for i in range(0, xq.size, m)])  # (k, m) layout
        assert np.array_equal(whole, parts)
    assert "_packed" in p.__dict__

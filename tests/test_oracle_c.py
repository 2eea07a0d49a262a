"""The OpenMP C restatement (oracle/interpax_c.c, the CPU baseline) agrees with the NumPy
oracle to rounding.  CPU only."""

import numpy as np
import pytest

from oracle import interpax_c as OC
from oracle import interpax_np as O


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_c_oracle_matches_numpy_oracle(ndim):
    rng = np.random.default_rng(ndim)
    ns = [9, 7, 8][:ndim]
    knots = [np.sort(rng.uniform(0, 1, n)) for n in ns]
    f = rng.normal(size=ns)
    names = O.TABLES_3D if ndim == 3 else O.TABLES_2D if ndim == 2 else ("f", "fx")
    tabs = {"f": f}
    chain = {1: (("fx", "f", 0),), 2: (("fx", "f", 0), ("fy", "f", 1), ("fxy", "fx", 1)),
             3: (("fx", "f", 0), ("fy", "f", 1), ("fz", "f", 2), ("fxy", "fx", 1), ("fxz", "fx", 2),
                 ("fyz", "fy", 2), ("fxyz", "fxy", 2))}[ndim]
    for name, src, ax in chain:
        tabs[name] = O.approx_df(knots[ax], tabs[src], "cubic", ax)
    qs = [rng.uniform(-0.1, 1.1, 2000) for _ in range(ndim)]
    fun = (O.interp1d, O.interp2d, O.interp3d)[ndim - 1]
    for deriv in (0, 1, 2, 3, 4):
        for extrap in (False, True, (True, 0.5)):
            want = fun(*qs, *knots, f, "cubic", deriv, extrap, None, **{n: tabs[n] for n in names[1:]})
            got = OC.interp_cubic(qs, knots, [tabs[n] for n in names], deriv, extrap)
            fin = np.isfinite(want)
            assert np.array_equal(np.isnan(got), np.isnan(want))
            np.testing.assert_allclose(got[fin], want[fin], rtol=1e-12, atol=1e-12 * np.abs(want[fin]).max())
    assert OC.num_threads() >= 1

"""Smallest run that touches every kernel family once, for compute-sanitizer --tool memcheck:
tile kernel (1/2/3-D, f32/f64, periodic), rows kernel, generic kernel (SoA, linear, nearest,
batch), slopes (all methods incl. windowed cubic2), periodic knots, pad, pack, presort, vjp."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402

rng = np.random.default_rng(0)
for dt in (np.float64, np.float32):
    for ndim in (1, 2, 3):
        ns = [23, 17, 13][:ndim]
        knots = [np.sort(rng.uniform(0, 2, n)).astype(dt) for n in ns]
        f = rng.normal(size=ns).astype(dt)
        fb = rng.normal(size=(*ns, 3)).astype(dt)
        qs = [rng.uniform(-0.3, 2.3, 777).astype(dt) for _ in range(ndim)]
        qs[0][3] = np.nan
        cls = (ipx.Interpolator1D, ipx.Interpolator2D, ipx.Interpolator3D)[ndim - 1]
        fun = (ipx.interp1d, ipx.interp2d, ipx.interp3d)[ndim - 1]
        per = 2.1 if ndim == 1 else (2.1,) + (None,) * (ndim - 1)
        for method in ("cubic", "cubic2", "cardinal", "monotonic", "monotonic-0", "akima", "linear", "nearest"):
            cls(*knots, f, method, False, per)(*qs)
            cls(*knots, fb, method, True)(*qs)
            fun(*qs, *knots, f, method, 0, (True, 0.5), per)
        cls(*knots, f, "cubic", True)(*[np.sort(q) for q in qs], presort=True)
        ipx.vjp_tables(qs, knots, f.shape, rng.normal(size=777).astype(dt), "cubic", 0, True, per)
    x = np.sort(rng.uniform(0, 1, 1500)).astype(dt)
    ipx.approx_df(x, rng.normal(size=(1500, 2)).astype(dt), "cubic2", 0)
    wide = rng.normal(size=(40, 70)).astype(dt)
    xk = np.linspace(0, 1, 40).astype(dt)
    ipx.Interpolator1D(xk, wide, "cubic")(rng.uniform(-0.1, 1.1, 500).astype(dt))
    ipx.interp1d(rng.uniform(-0.1, 1.1, 500).astype(dt), xk, wide, "linear")
print("sanitize_small done")

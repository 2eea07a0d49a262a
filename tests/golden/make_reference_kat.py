"""Writes tests/golden/reference_kat.json: the known-answer vectors that the reference's
own tests hold for the slope stencils and the 1-D evaluation path.

The reference (f0uriest/interpax) cannot be imported in this environment (no jax), so
these are not outputs of the reference run here; they are the literal golden numbers in
the reference's test files, transcribed with the file:line they come from, plus the
hash of the coefficient matrices read from /root/reference/interpax/_coefs.py (that file
needs only numpy, so it *is* loaded when the script runs in the build container).

Run:  python tests/golden/make_reference_kat.py
"""

import hashlib
import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

kat = {}

# tests/test_scipy.py:68-91  Akima1DInterpolator golden values
kat["akima_eval"] = {
    "cite": "tests/test_scipy.py:68-91",
    "x": list(np.arange(0.0, 11.0)),
    "y": [0.0, 2.0, 1.0, 3.0, 2.0, 6.0, 5.5, 5.5, 2.7, 5.1, 3.0],
    "xi": [0.0, 0.5, 1.0, 1.5, 2.5, 3.5, 4.5, 5.1, 6.5, 7.2, 8.6, 9.9, 10.0],
    "yi": [
        0.0,
        1.375,
        2.0,
        1.5,
        1.953125,
        2.484375,
        4.1363636363636366866103344,
        5.9803623910336236590978842,
        5.5067291516462386624652936,
        5.2031367459745245795943447,
        4.1796554159017080820603951,
        3.4110386597938129327189927,
        3.0,
    ],
    "rtol": 1e-7,
}

# tests/test_scipy.py:157-164  degenerate Akima (issue 5683)
kat["akima_degenerate"] = {
    "cite": "tests/test_scipy.py:157-164",
    "x": [0.0, 1.0, 2.0],
    "x_eval": [0.5, 1.5],
}

# tests/test_scipy.py:653-686  PCHIP, NAG e01bec example
kat["pchip_nag"] = {
    "cite": "tests/test_scipy.py:653-686",
    "x": [7.99, 8.09, 8.19, 8.70, 9.20, 10.00, 12.00, 15.00, 20.00],
    "y": [0.0, 0.27643e-4, 0.43750e-1, 0.16918, 0.46943, 0.94374, 0.99864, 0.99992, 0.99999],
    "xi": [7.99, 9.191, 10.392, 11.593, 12.794, 13.995, 15.196, 16.397, 17.598, 18.799, 20.0],
    "yi": [0.0, 0.4640, 0.9645, 0.9965, 0.9992, 0.9998, 0.9999, 1.0, 1.0, 1.0, 1.0],
    "atol": 5e-5,
}

# tests/test_scipy.py:637-651  PCHIP integer-cast regression data
kat["pchip_cast"] = {
    "cite": "tests/test_scipy.py:637-651",
    "x": [0, 4, 12, 27, 47, 60, 79, 87, 99, 100],
    "y": [-33, -33, -19, -2, 12, 26, 38, 45, 53, 55],
}

# tests/test_scipy.py:688-696  PCHIP end slopes must be non-zero
kat["pchip_endslopes"] = {
    "cite": "tests/test_scipy.py:688-696",
    "x": [0.0, 0.1, 0.25, 0.35],
    "y1": [279.35, 0.5e3, 1.0e3, 2.5e3],
    "y2": [279.35, 2.5e3, 1.50e3, 1.0e3],
}

# tests/test_scipy.py:710-717  two knots -> straight line
kat["pchip_two_points"] = {"cite": "tests/test_scipy.py:710-717", "x": [0, 1], "y": [0, 2]}

# tests/test_scipy.py:719-732  three knots
kat["pchip_three_points"] = {
    "cite": "tests/test_scipy.py:719-732",
    "x": [1, 2, 3],
    "y": [4, 5, 6],
    "xi": [0.5],
    "d0": [3.5],
    "d1": [1.0],
}

# tests/test_interpolate.py:136-149  monotonic sign test
kat["monotonic_heaviside"] = {"cite": "tests/test_interpolate.py:136-149"}

# coefficient matrices, interpax/_coefs.py:8-163
ref = "/root/reference/interpax/_coefs.py"
if os.path.exists(ref):
    spec = importlib.util.spec_from_file_location("_ref_coefs", ref)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    coefs = {}
    for name in ("A_CUBIC", "A_BICUBIC", "A_TRICUBIC"):
        a = np.ascontiguousarray(getattr(mod, name).astype(np.int64))
        coefs[name] = {
            "shape": list(a.shape),
            "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
            "nnz": int(np.count_nonzero(a)),
            "absmax": int(np.abs(a).max()),
            "rowsum": [int(v) for v in a.sum(axis=1)],
        }
    kat["coefs"] = {"cite": "interpax/_coefs.py:8-163", **coefs}
else:  # keep what is already on disk
    with open(os.path.join(HERE, "reference_kat.json")) as fh:
        kat["coefs"] = json.load(fh)["coefs"]

with open(os.path.join(HERE, "reference_kat.json"), "w") as fh:
    json.dump(kat, fh, indent=1)
print("wrote", os.path.join(HERE, "reference_kat.json"))

"""ctypes wrapper of oracle/interpax_c.c (OpenMP C restatement of the reference's cubic
evaluation).  TEST INFRASTRUCTURE ONLY -- see the header of interpax_c.c."""

import ctypes as C
import os
import subprocess

import numpy as np

from . import interpax_np as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_c.so")


class Extrap(C.Structure):
    _fields_ = [("keep_lo", C.c_int), ("keep_hi", C.c_int), ("fill_lo", C.c_double), ("fill_hi", C.c_double)]


_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "interpax_c.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.ipxo_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(load().ipxo_num_threads())


def use_all_threads() -> int:
    """Use every host thread this process may run on (ignores OMP_NUM_THREADS=1 from torchrun)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    load().ipxo_set_threads(C.c_int(n))
    return num_threads()


def _ex(extrap, n):
    e6 = O._parse_extrap(extrap, n)
    arr = (Extrap * n)()
    for ax in range(n):
        for side, name in ((0, "lo"), (1, "hi")):
            v = e6[2 * ax + side]
            keep = 1 if (O.isbool(v) and bool(v)) else 0
            fill = float("nan") if O.isbool(v) else float(v)
            setattr(arr[ax], "keep_" + name, keep)
            setattr(arr[ax], "fill_" + name, fill)
    return arr


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def interp_cubic(qs, knots, tables, derivative=0, extrap=False):
    """Cubic-family evaluation with all 2^d tables given (float64, scalar-valued f,
    non-periodic: pad first for periodic axes).  qs / knots: tuples of d arrays."""
    lib = load()
    d = len(qs)
    qs = [np.ascontiguousarray(q, dtype=np.float64) for q in qs]
    knots = [np.ascontiguousarray(k, dtype=np.float64) for k in knots]
    tables = [np.ascontiguousarray(t, dtype=np.float64) for t in tables]
    nq = qs[0].size
    out = np.empty(nq, dtype=np.float64)
    der = (C.c_int * d)(*[int(v) for v in O._parse_ndarg(derivative, d)])
    ex = _ex(extrap, d)
    tabs = (C.c_void_p * len(tables))(*[t.ctypes.data for t in tables])
    if d == 3:
        lib.ipxo_interp3d_cubic(_dp(qs[0]), _dp(qs[1]), _dp(qs[2]), C.c_int64(nq), _dp(knots[0]),
                                C.c_int(knots[0].size), _dp(knots[1]), C.c_int(knots[1].size),
                                _dp(knots[2]), C.c_int(knots[2].size), tabs, der, ex, _dp(out))
    elif d == 2:
        lib.ipxo_interp2d_cubic(_dp(qs[0]), _dp(qs[1]), C.c_int64(nq), _dp(knots[0]),
                                C.c_int(knots[0].size), _dp(knots[1]), C.c_int(knots[1].size), tabs,
                                der, ex, _dp(out))
    else:
        lib.ipxo_interp1d_cubic(_dp(qs[0]), C.c_int64(nq), _dp(knots[0]), C.c_int(knots[0].size), tabs,
                                der, ex, _dp(out))
    return out

"""CPU-side checks of the drop-in boundary: libipx.so loads, exports every symbol that
include/ipx.h declares, the ctypes mirror of the structs has the header's layout, and the
host wrapper rejects bad calls with the reference's errors before touching a device.
No kernels are launched here."""

import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import interpax_b200 as ipx
from interpax_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ipx.h")


def header_symbols():
    src = open(HEADER).read()
    names = set(re.findall(r"\b(ipx_[a-z0-9_]+)\s*\(", src))
    names -= {n for n in names if "##" in n}
    # macro-generated evaluation entry points
    out = set()
    for n in names:
        if n.endswith("_"):
            continue
        out.add(n)
    for d in (1, 2, 3):
        for t in ("f32", "f64"):
            out.add(f"ipx_eval{d}d_{t}")
    # macro-generated derivative entry points (IPX_DECL_GRAD)
    for t in ("f32", "f64"):
        out.add(f"ipx_vjp_stencil_{t}")
        out.add(f"ipx_jvp_stencil_{t}")
        # (IPX_DECL_GRAD_TABLES)
        out.add(f"ipx_vjp_knots_tables_{t}")
        out.add(f"ipx_slopes_knots_vjp_{t}")
        out.add(f"ipx_jvp_knots_tables_{t}")
        out.add(f"ipx_slopes_jvp_rows_{t}")
        out.add(f"ipx_cubic2_solve_{t}")
    return {n for n in out if not n.startswith("ipx_eval") or re.fullmatch(r"ipx_eval([123]d|_stencil|_stencil_perm)_f(32|64)", n)}


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_symbols()
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ipx_version() == 100
    assert b"sm_100a" in lib.ipx_build_info()


def test_library_is_sm100a_sass():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layout_matches_header():
    # compile a tiny C program against the header and compare sizeof/offsetof
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "ipx.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ipx_axis_params), sizeof(ipx_params),
             offsetof(ipx_axis_params, fill_lo), offsetof(ipx_axis_params, period),
             sizeof(ipx_slope_opts), offsetof(ipx_slope_opts, bc_start),
             offsetof(ipx_slope_opts, bc_start_value), offsetof(ipx_slope_opts, bc_end_value));
      return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        cpath = os.path.join(td, "t.c")
        open(cpath, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), cpath, "-o", exe], check=True)
        vals = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert vals == [
        C.sizeof(_lib.AxisParams), C.sizeof(_lib.Params), _lib.AxisParams.fill_lo.offset,
        _lib.AxisParams.period.offset, C.sizeof(_lib.SlopeOpts), _lib.SlopeOpts.bc_start.offset,
        _lib.SlopeOpts.bc_start_value.offset, _lib.SlopeOpts.bc_end_value.offset,
    ]


def test_argument_errors_match_reference_without_a_device():
    x = np.arange(4.0)
    with pytest.raises(ValueError, match="x and f must be arrays of equal length"):
        ipx.interp1d(np.zeros(3), x, np.arange(5.0))
    with pytest.raises(ValueError, match="unknown method foo"):
        ipx.interp1d(np.zeros(3), x, x, method="foo")
    with pytest.raises(ValueError, match="y and f must be arrays of equal length"):
        ipx.interp2d(0.0, 0.0, x, np.arange(3.0), np.zeros((4, 4)))
    with pytest.raises(ValueError, match="z and f must be arrays of equal length"):
        ipx.Interpolator3D(x, x, np.arange(3.0), np.zeros((4, 4, 4)))
    with pytest.raises(ValueError, match="unknown method"):
        ipx.Interpolator2D(x, x, np.zeros((4, 4)), method="quintic")


def test_extrap_and_ndarg_parsing():
    from interpax_b200 import api
    assert api._parse_extrap(False, 2) == (False,) * 4
    assert api._parse_extrap(1.5, 1) == (1.5, 1.5)
    assert api._parse_extrap((True, 0.0), 3) == (True, 0.0) * 3
    assert api._parse_extrap(((True, 1.0), (False, 2.0)), 2) == (True, 1.0, False, 2.0)
    with pytest.raises(ValueError, match="extrap should either be a scalar"):
        api._parse_extrap(((1, 2, 3),), 2)
    assert api._parse_ndarg(2, 3) == (2, 2, 2)
    assert api._parse_ndarg((1, 0), 2) == (1, 0)
    assert api._keep_fill(True) == (1, 0.0)
    k, v = api._keep_fill(False)
    assert k == 0 and np.isnan(v)
    assert api._keep_fill(2.5) == (0, 2.5)
    # dtype promotion: Python scalars are weak, integers become float64 (utils.py:42-50)
    assert api._promote([np.zeros(2, np.float32), 0.0]) == np.float32
    assert api._promote([np.zeros(2, np.float32), np.zeros(2, np.float64)]) == np.float64
    assert api._promote([np.arange(3)]) == np.float64
    assert api._promote([1, 2.0]) == np.float64


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ipx.interp1d(np.zeros(3), np.arange(4.0), np.arange(4.0))
    # the product package never imports the oracle
    import sys
    pkg = os.path.join(ROOT, "interpax_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no oracle", ""), fn


def test_header_constants_match_the_python_mirror():
    """Every integer `#define IPX_*` of include/ipx.h that interpax_b200/_lib.py mirrors has the
    same value there (method classes, layouts, slope codes, boundary kinds, error codes)."""
    src = open(HEADER).read()
    defines = {m.group(1): int(m.group(2), 0)
               for m in re.finditer(r"^#define\s+(IPX_[A-Z0-9_]+)\s+\(?(-?(?:0x)?[0-9a-fA-F]+)\)?\s*(?:/\*.*)?$",
                                    src, re.M)}
    assert "IPX_SLOPE_AKIMA_COMPLEX" in defines and "IPX_OK" in defines
    mirrored = [n for n in defines if hasattr(_lib, n) and isinstance(getattr(_lib, n), int)]
    assert len(mirrored) >= 15, mirrored
    for n in mirrored:
        assert getattr(_lib, n) == defines[n], n

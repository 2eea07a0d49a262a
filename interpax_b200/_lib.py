"""ctypes binding of libipx.so -- the C ABI declared in include/ipx.h.

There is no fallback: if the library cannot be loaded the import raises, and every
entry point raises on a non-zero return code.
"""

from __future__ import annotations

import ctypes as C
import os

from . import _build

# ---- constants mirrored from include/ipx.h -------------------------------------------
IPX_OK, IPX_EINVAL, IPX_EUNSUPPORTED, IPX_ELAUNCH = 0, -1, -2, -3
IPX_NEAREST, IPX_LINEAR, IPX_CUBIC = 0, 1, 2
IPX_SOA, IPX_AOS = 0, 1
IPX_SLOPE_CUBIC, IPX_SLOPE_CUBIC2, IPX_SLOPE_CARDINAL = 0, 1, 2
IPX_SLOPE_MONOTONIC, IPX_SLOPE_MONOTONIC0, IPX_SLOPE_AKIMA, IPX_SLOPE_ZERO = 3, 4, 5, 6
IPX_SLOPE_AKIMA_COMPLEX = 7
IPX_PPOLY_NAN, IPX_PPOLY_EXTRAPOLATE, IPX_PPOLY_PERIODIC = 0, 1, 2
IPX_BC_NOT_A_KNOT, IPX_BC_FIRST, IPX_BC_SECOND = 0, 1, 2
IPX_STENCIL_AUTO, IPX_STENCIL_SPARSE, IPX_STENCIL_DENSE, IPX_STENCIL_COLUMN = 0, 1, 2, 3


def IPX_PERIODIC(axis: int) -> int:
    return 1 << axis


def IPX_WRAP_ONLY(axis: int) -> int:
    return 1 << (4 + axis)


def IPX_PAD_FIRST(axis: int) -> int:
    return 1 << (4 + axis)


class AxisParams(C.Structure):
    _fields_ = [
        ("derivative", C.c_int32),
        ("keep_lo", C.c_int32),
        ("keep_hi", C.c_int32),
        ("reserved", C.c_int32),
        ("fill_lo", C.c_double),
        ("fill_hi", C.c_double),
        ("period", C.c_double),
    ]


class Params(C.Structure):
    _fields_ = [("axis", AxisParams * 3)]


class SlopeOpts(C.Structure):
    _fields_ = [
        ("c", C.c_double),
        ("bc_start", C.c_int32),
        ("bc_end", C.c_int32),
        ("bc_start_value", C.c_void_p),
        ("bc_end_value", C.c_void_p),
    ]


# every symbol include/ipx.h declares; tests/test_abi.py checks the .so exports them all
SYMBOLS = (
    ["ipx_version", "ipx_build_info", "ipx_slopes_workspace_bytes", "ipx_slopes_vjp_workspace_bytes", "ipx_cell_order_workspace_bytes",
     "ipx_cell_sort_workspace_bytes", "ipx_halo_table_elems", "ipx_launch_count"]
    + [f"ipx_eval{d}d_{t}" for d in (1, 2, 3) for t in ("f32", "f64")]
    + [
        f"ipx_{name}_{t}"
        for name in ("bin_index", "slopes", "slopes_vjp", "slopes_pack", "periodic_knots", "pad_periodic", "pack", "cell_order", "cell_sort", "cell_order_records", "eval_stencil_perm", "gather", "halo_table", "eval_stencil",
                     "gather_multi", "scatter", "vjp_tables", "vjp_stencil", "jvp_stencil", "vjp_knots_tables", "slopes_knots_vjp", "jvp_knots_tables", "slopes_jvp_rows", "cubic2_solve", "ppoly_eval", "ppoly_pack", "ppoly_from_hermite",
                     "ppoly_derivative", "ppoly_antiderivative")
        for t in ("f32", "f64")
    ]
)


class IpxError(RuntimeError):
    pass


_ERR = {IPX_EINVAL: "IPX_EINVAL (bad argument)", IPX_EUNSUPPORTED: "IPX_EUNSUPPORTED",
        IPX_ELAUNCH: "IPX_ELAUNCH (CUDA launch failed)"}


def check(rc: int, what: str) -> None:
    if rc != IPX_OK:
        raise IpxError(f"{what} failed: {_ERR.get(rc, rc)}")


def _declare(lib: C.CDLL) -> None:
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.ipx_version.restype = C.c_int
    lib.ipx_version.argtypes = []
    lib.ipx_build_info.restype = C.c_char_p
    lib.ipx_build_info.argtypes = []
    lib.ipx_slopes_workspace_bytes.restype = C.c_size_t
    lib.ipx_slopes_workspace_bytes.argtypes = [i32, i32, i32]
    lib.ipx_slopes_vjp_workspace_bytes.restype = C.c_size_t
    lib.ipx_slopes_vjp_workspace_bytes.argtypes = [i32, i32, i64, i64, i32]
    lib.ipx_launch_count.restype = C.c_uint64
    lib.ipx_launch_count.argtypes = []
    lib.ipx_halo_table_elems.restype = C.c_int64
    lib.ipx_halo_table_elems.argtypes = [i32, C.POINTER(i32), i32, i32]
    lib.ipx_cell_order_workspace_bytes.restype = C.c_size_t
    lib.ipx_cell_order_workspace_bytes.argtypes = [i64]
    lib.ipx_cell_sort_workspace_bytes.restype = C.c_size_t
    lib.ipx_cell_sort_workspace_bytes.argtypes = [i32, i64, i32]
    tail = [C.POINTER(vp), i32, i64, i32, i32, C.POINTER(Params), i32, vp, vp, vp]
    for t in ("f32", "f64"):
        f = getattr(lib, f"ipx_eval1d_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i64, vp, vp, i32] + tail
        f = getattr(lib, f"ipx_eval2d_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, vp, i64, vp, vp, vp, vp, i32, i32] + tail
        f = getattr(lib, f"ipx_eval3d_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i32, i32, i32] + tail
        f = getattr(lib, f"ipx_bin_index_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, vp, i64, vp, vp]
        f = getattr(lib, f"ipx_halo_table_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, vp, C.POINTER(i32), i32, C.POINTER(vp), vp, vp]
        f = getattr(lib, f"ipx_eval_stencil_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), vp, vp, i32,
                      C.c_double, i32, C.POINTER(Params), i32, vp, vp, i32, vp, vp]
        f = getattr(lib, f"ipx_slopes_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, vp, i64, i64, i32, C.POINTER(SlopeOpts), vp, vp, vp]
        f = getattr(lib, f"ipx_vjp_knots_tables_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), i32,
                      C.POINTER(Params), i32, vp, C.POINTER(vp), vp]
        f = getattr(lib, f"ipx_jvp_knots_tables_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), i32,
                      C.POINTER(Params), i32, C.POINTER(vp), vp, vp]
        f = getattr(lib, f"ipx_slopes_jvp_rows_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, vp, vp, vp, vp, i64, i64, i32, C.POINTER(SlopeOpts), vp, vp]
        f = getattr(lib, f"ipx_cubic2_solve_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, i32, i32, vp, i64, i64, vp, vp]
        f = getattr(lib, f"ipx_slopes_knots_vjp_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, vp, vp, vp, i64, i64, i32, C.POINTER(SlopeOpts), vp, vp, vp]
        f = getattr(lib, f"ipx_slopes_vjp_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, vp, i64, i64, i32, C.POINTER(SlopeOpts), i32, vp, vp, vp]
        f = getattr(lib, f"ipx_periodic_knots_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, C.c_double, vp, vp, vp, vp]
        f = getattr(lib, f"ipx_pad_periodic_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i64, i32, i64, vp, vp, vp]
        f = getattr(lib, f"ipx_pack_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, vp, vp]
        f = getattr(lib, f"ipx_slopes_pack_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), C.POINTER(i32), vp, i64, i32, C.POINTER(SlopeOpts), vp, vp]
        f = getattr(lib, f"ipx_cell_order_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(i32), i32, C.POINTER(Params), i32,
                      vp, vp, vp]
        f = getattr(lib, f"ipx_cell_order_records_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(i32), i32, C.POINTER(Params), i32,
                      vp, vp, vp, vp]
        f = getattr(lib, f"ipx_eval_stencil_perm_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), vp, i32,
                      C.c_double, i32, C.POINTER(Params), i32, vp, vp, vp, vp]
        f = getattr(lib, f"ipx_cell_sort_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(i32), i32, C.POINTER(Params), i32,
                      vp, C.POINTER(vp), vp, vp]
        f = getattr(lib, f"ipx_ppoly_eval_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, i32, i32, i64, vp, vp, i64, i32, i32, vp, vp, vp]
        f = getattr(lib, f"ipx_ppoly_pack_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, i32, vp, vp]
        f = getattr(lib, f"ipx_ppoly_from_hermite_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, vp, vp, i64, vp, vp]
        f = getattr(lib, f"ipx_ppoly_derivative_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, i32, i64, i32, vp, vp]
        f = getattr(lib, f"ipx_ppoly_antiderivative_{t}")
        f.restype = C.c_int
        f.argtypes = [vp, i32, i32, i64, vp, vp, vp]
        f = getattr(lib, f"ipx_vjp_tables_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), i64, i32, i32,
                      C.POINTER(Params), i32, vp, C.POINTER(vp), vp]
        f = getattr(lib, f"ipx_vjp_stencil_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), vp, i32,
                      C.c_double, i32, C.POINTER(Params), i32, vp, C.POINTER(vp), vp, vp]
        f = getattr(lib, f"ipx_jvp_stencil_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), vp, i32,
                      C.c_double, i32, C.POINTER(Params), i32, C.POINTER(vp), vp, vp, vp]
        f = getattr(lib, f"ipx_gather_multi_{t}")
        f.restype = C.c_int
        f.argtypes = [i32, C.POINTER(vp), vp, i64, C.POINTER(vp), vp]
        for name in ("gather", "scatter"):
            f = getattr(lib, f"ipx_{name}_{t}")
            f.restype = C.c_int
            f.argtypes = [vp, vp, i64, i64, vp, vp]


_LIB = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if the sources are newer and nvcc is present) libipx.so."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("IPX_LIB")  # experiments: another build of the same sources
    if path is None:
        path = _build.LIB
        if _build.is_stale() and _build.find_nvcc() is not None:
            _build.build()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing and nvcc is not available to build it: the interpax_b200 "
            "query path is CUDA-only (no CPU fallback). Run `python -m interpax_b200._build`."
        )
    lib = C.CDLL(path)
    _declare(lib)
    _LIB = lib
    return lib

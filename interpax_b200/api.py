"""Host side of the interpax query path on B200: the reference's call signatures on top of
the C ABI (include/ipx.h, libipx.so).

Mirrors (reference repo paths):
  interp1d / interp2d / interp3d            interpax/_spline.py:444-1105
  Interpolator1D / 2D / 3D                  interpax/_spline.py:48-441
  approx_df                                 interpax/_fd_derivs.py:9-76
with the same argument meaning, dtype promotion (utils.py:42-50, _spline.py:503-506),
output shapes and ValueError texts.  All arithmetic happens in the CUDA kernels; this
file only parses arguments and moves arrays.  PyTorch is used for device memory and
streams, nothing else.  Inputs may be NumPy arrays / Python scalars (results come back
as NumPy arrays: host -> device -> host) or torch tensors (results stay on the device).

There is no CPU fallback: without a CUDA device or without libipx.so every call raises.
"""

from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Any

import numpy as np
import torch

from . import _lib as L

CUBIC_METHODS = ("cubic", "cubic2", "cardinal", "catmull-rom", "akima", "monotonic", "monotonic-0")
OTHER_METHODS = ("nearest", "linear")
METHODS_1D = METHODS_2D = METHODS_3D = CUBIC_METHODS + OTHER_METHODS

# the reference's stacking order of the slope tables (_spline.py:581-584, 779-783, 1070-1078)
TABLES = {
    1: ("f", "fx"),
    2: ("f", "fx", "fy", "fxy"),
    3: ("f", "fx", "fy", "fz", "fxy", "fxz", "fyz", "fxyz"),
}
# chain used to fill in missing tables: name -> (source table, axis)
# (_spline.py:119-122, 237-244, 380-393, 572-573, 759-764, 1027-1040)
CHAIN = {
    1: (("fx", "f", 0),),
    2: (("fx", "f", 0), ("fy", "f", 1), ("fxy", "fx", 1)),
    3: (
        ("fx", "f", 0), ("fy", "f", 1), ("fz", "f", 2), ("fxy", "fx", 1),
        ("fxz", "fx", 2), ("fyz", "fy", 2), ("fxyz", "fxy", 2),
    ),
}
_SLOPE_CODE = {
    "cubic": L.IPX_SLOPE_CUBIC, "cubic2": L.IPX_SLOPE_CUBIC2, "cardinal": L.IPX_SLOPE_CARDINAL,
    "catmull-rom": L.IPX_SLOPE_CARDINAL, "monotonic": L.IPX_SLOPE_MONOTONIC,
    "monotonic-0": L.IPX_SLOPE_MONOTONIC0, "akima": L.IPX_SLOPE_AKIMA,
    "nearest": L.IPX_SLOPE_ZERO, "linear": L.IPX_SLOPE_ZERO,
}
_NP2T = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
         np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}
_T2NP = {
    torch.float16: np.dtype(np.float16), torch.bfloat16: np.dtype(np.float32),
    torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64),
    torch.complex64: np.dtype(np.complex64), torch.complex128: np.dtype(np.complex128),
    torch.int8: np.dtype(np.int8), torch.int16: np.dtype(np.int16), torch.int32: np.dtype(np.int32),
    torch.int64: np.dtype(np.int64), torch.uint8: np.dtype(np.uint8), torch.bool: np.dtype(np.bool_),
}


# --------------------------------------------------------------------------- small helpers
def isbool(x: Any) -> bool:
    """utils.py:17-19."""
    return isinstance(x, bool) or (hasattr(x, "dtype") and (x.dtype == bool or x.dtype == torch.bool))


def errorif(cond, err=ValueError, msg=""):
    """utils.py:22-31."""
    if cond:
        raise err(msg)


_have_cuda = None
_devices: dict = {}


def _device() -> torch.device:
    global _have_cuda
    if _have_cuda is None:
        _have_cuda = bool(torch.cuda.is_available())
    if not _have_cuda:
        raise RuntimeError(
            "interpax_b200 needs a CUDA device: the query path is hand-written CUDA behind "
            "libipx.so and has no CPU fallback"
        )
    idx = torch.cuda.current_device()
    dev = _devices.get(idx)
    if dev is None:
        dev = _devices[idx] = torch.device("cuda", idx)
    return dev


try:  # the raw handle of the current stream without building a torch.cuda.Stream object per call
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    _raw_stream = None


def _stream() -> int:
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _is_pyscalar(a) -> bool:
    return isinstance(a, (bool, int, float, complex))


def _np_dtype_of(a) -> np.dtype:
    if isinstance(a, torch.Tensor):
        return _T2NP[a.dtype]
    return np.asarray(a).dtype


def _inexact(dt: np.dtype) -> np.dtype:
    """asarray_inexact (utils.py:42-50): integer / bool inputs become the default float."""
    if not np.issubdtype(dt, np.inexact):
        return np.dtype(np.float64)
    if dt == np.float16:
        return np.dtype(np.float32)
    return dt


def _promote(items) -> np.dtype:
    """result_type with Python scalars weakly typed, as in JAX (utils.py:45-46)."""
    parts = [a if _is_pyscalar(a) else _inexact(_np_dtype_of(a)) for a in items]
    dt = np.result_type(*parts)
    return _inexact(np.dtype(dt))


def _to_dev(a, np_dt: np.dtype, dev: torch.device) -> torch.Tensor:
    tdt = _NP2T[np.dtype(np_dt)]
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=tdt).contiguous()
    arr = np.array(a, dtype=np_dt, order="C")  # (np.ascontiguousarray would promote 0-d to 1-d)
    return torch.from_numpy(arr).to(dev)


def _parse_ndarg(arg, n):
    """_spline.py:1154-1161."""
    try:
        k = len(arg)
    except TypeError:
        arg = tuple(arg for _ in range(n))
        k = n
    assert k == n, "got too many args"
    return tuple(arg)


def _isscalar(v) -> bool:
    return np.isscalar(v) or (isinstance(v, (np.ndarray, torch.Tensor)) and v.ndim == 0)


def _parse_extrap(extrap, n):
    """_spline.py:1164-1177."""
    if isbool(extrap):
        return tuple(extrap for _ in range(2 * n))
    elif _isscalar(extrap):
        return tuple(extrap for _ in range(2 * n))
    elif len(extrap) == 2 and _isscalar(extrap[0]):
        return tuple(e for _ in range(n) for e in extrap)
    elif len(extrap) == n and all(len(extrap[i]) == 2 for i in range(n)):
        return tuple(eij for ei in extrap for eij in ei)
    raise ValueError(
        "extrap should either be a scalar, 2 element sequence (lo, hi), "
        + "or a sequence with 2 elements for each dimension"
    )


def _keep_fill(e):
    """One side of `extrap` -> (keep flag, fill value) of ipx_axis_params (_spline.py:1190-1220)."""
    if isbool(e):
        return (1, 0.0) if bool(e) else (0, math.nan)
    return 0, float(e)


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def _ptr_array(ts):
    arr = (C.c_void_p * len(ts))()
    for k, t in enumerate(ts):
        arr[k] = None if t is None else t.data_ptr()
    return arr


def _suffix(t: torch.Tensor) -> str:
    return "f32" if t.dtype == torch.float32 else "f64"


# --------------------------------------------------------------------------- device-side ops
def _slopes_dev(x: torch.Tensor, f: torch.Tensor, method: str, axis: int, kwargs: dict,
                cplx: bool = False) -> torch.Tensor:
    """approx_df on device tensors (real dtype, C-contiguous).  ``cplx``: ``f`` is complex data
    viewed as a trailing (re, im) pair; only akima needs to know (its weights are moduli of
    complex differences, _fd_derivs.py:414), every other method acts on the parts separately."""
    errorif(method not in _SLOPE_CODE, ValueError, f"got unknown method {method}")
    lib = L.load()
    axis = axis % f.ndim
    n = f.shape[axis]
    errorif(x.ndim != 1 or x.shape[0] != n, ValueError, "x and f must be arrays of equal length")
    outer = int(np.prod(f.shape[:axis], dtype=np.int64))
    inner = int(np.prod(f.shape[axis + 1:], dtype=np.int64))
    code = _SLOPE_CODE[method]
    if cplx and method == "akima":
        errorif(axis == f.ndim - 1, ValueError, "the (re, im) pair axis cannot be differentiated")
        code = L.IPX_SLOPE_AKIMA_COMPLEX
    opts, keep = _slope_opts_values(f, method, axis, kwargs, cplx)
    out = torch.empty_like(f)
    elem = f.element_size()
    wbytes = lib.ipx_slopes_workspace_bytes(code, n, elem)
    ws = torch.empty(max(int(wbytes), 16), dtype=torch.uint8, device=f.device)
    fn = getattr(lib, f"ipx_slopes_{_suffix(f)}")
    L.check(
        fn(_ptr(x), n, _ptr(f), outer, inner, code, C.byref(opts), _ptr(out), _ptr(ws), _stream()),
        "ipx_slopes",
    )
    return out


def _slope_opts_values(f: torch.Tensor, method: str, axis: int, kwargs: dict, cplx: bool = False):
    """ipx_slope_opts of a forward call (tension, boundary kinds and VALUES) + the tensors it points to."""
    opts = L.SlopeOpts()
    opts.c = float(kwargs.get("c", 0)) if method == "cardinal" else 0.0
    keep = []
    if method == "cubic2":
        # (complex data arrives as a trailing (re, im) pair: the boundary values have the complex shape)
        expect = f.shape[:axis] + f.shape[axis + 1:]
        bcs = _validate_bc(kwargs.get("bc_type", "not-a-knot"), expect[:-1] if cplx else expect)
        kinds, vals = [], []
        for kind, val in bcs:
            kinds.append(kind)
            if val is None:
                vals.append(None)
            else:
                if cplx:
                    cdt = np.complex128 if f.dtype == torch.float64 else np.complex64
                    v = torch.view_as_real(_to_dev(val, cdt, f.device).contiguous()).reshape(-1)
                else:
                    v = _to_dev(val, _T2NP[f.dtype], f.device).reshape(-1)
                keep.append(v)
                vals.append(v.data_ptr())
        opts.bc_start, opts.bc_end = kinds
        opts.bc_start_value, opts.bc_end_value = vals
    return opts, keep


def _validate_bc(bc_type, expected_shape):
    """_fd_derivs.py:104-157 -> [(IPX_BC_*, value-or-None)] * 2."""
    if isinstance(bc_type, str):
        errorif(bc_type == "periodic", NotImplementedError)
        bc_type = (bc_type, bc_type)
    else:
        errorif(
            len(bc_type) != 2, ValueError,
            "`bc_type` must contain 2 elements to specify start and end conditions.",
        )
        errorif(
            any(isinstance(b, str) and b == "periodic" for b in bc_type), ValueError,
            "'periodic' `bc_type` is defined for both curve ends and cannot be used with other "
            "boundary conditions.",
        )
    out = []
    for bc in bc_type:
        if isinstance(bc, str):
            if bc == "clamped":
                out.append((L.IPX_BC_FIRST, None))
            elif bc == "natural":
                out.append((L.IPX_BC_SECOND, None))
            elif bc == "not-a-knot":
                out.append((L.IPX_BC_NOT_A_KNOT, None))
            else:
                raise ValueError(f"bc_type={bc} is not allowed.")
        else:
            try:
                order, value = bc
            except Exception as e:
                raise ValueError(
                    "A specified derivative value must be given in the form (order, value)."
                ) from e
            if order not in [1, 2]:
                raise ValueError("The specified derivative order must be 1 or 2.")
            shape = tuple(value.shape) if hasattr(value, "shape") else np.shape(value)
            if tuple(shape) != tuple(expected_shape):
                raise ValueError(
                    "`deriv_value` shape {} is not the expected one {}.".format(
                        tuple(shape), tuple(expected_shape)
                    )
                )
            out.append((L.IPX_BC_FIRST if order == 1 else L.IPX_BC_SECOND, value))
    return out


def _periodic_knots_dev(x: torch.Tensor, period: float):
    """Knot side of _make_periodic (_spline.py:1117-1122) -> (xpad[n+2], perm[n])."""
    lib = L.load()
    n = x.shape[0]
    xpad = torch.empty(n + 2, dtype=x.dtype, device=x.device)
    perm = torch.empty(n, dtype=torch.int32, device=x.device)
    ws = torch.empty((n + 4) * x.element_size(), dtype=torch.uint8, device=x.device)
    fn = getattr(lib, f"ipx_periodic_knots_{_suffix(x)}")
    L.check(fn(_ptr(x), n, abs(float(period)), _ptr(xpad), _ptr(perm), _ptr(ws), _stream()),
            "ipx_periodic_knots")
    return xpad, perm


def _pad_periodic_dev(t: torch.Tensor, axis: int, perm: torch.Tensor) -> torch.Tensor:
    """Materialised table side of _make_periodic (_spline.py:1123-1135)."""
    lib = L.load()
    n = t.shape[axis]
    outer = int(np.prod(t.shape[:axis], dtype=np.int64))
    inner = int(np.prod(t.shape[axis + 1:], dtype=np.int64))
    shape = list(t.shape)
    shape[axis] = n + 2
    out = torch.empty(shape, dtype=t.dtype, device=t.device)
    fn = getattr(lib, f"ipx_pad_periodic_{_suffix(t)}")
    L.check(fn(_ptr(t), outer, n, inner, _ptr(perm), _ptr(out), _stream()), "ipx_pad_periodic")
    return out


def _pack_dev(ndim: int, tabs: list[torch.Tensor]) -> torch.Tensor:
    """SoA -> AoS interleave; record layout in DESIGN.md."""
    lib = L.load()
    f = tabs[0]
    total = f.numel()
    out = torch.empty(tuple(f.shape) + (1 << ndim,), dtype=f.dtype, device=f.device)
    fn = getattr(lib, f"ipx_pack_{_suffix(f)}")
    L.check(fn(ndim, _ptr_array(tabs), total, _ptr(out), _stream()), "ipx_pack")
    return out


def _cell_order_dev(ndim, qs, knots, ns, periodic_mask, params) -> torch.Tensor:
    """Visiting order of the queries by table cell (stable argsort of the linear cell index)."""
    lib = L.load()
    nq = qs[0].numel()
    order = torch.empty(nq, dtype=torch.int32, device=qs[0].device)  # holds uint32 values
    ws = torch.empty(max(int(lib.ipx_cell_order_workspace_bytes(nq)), 256), dtype=torch.uint8,
                     device=qs[0].device)
    fn = getattr(lib, f"ipx_cell_order_{_suffix(qs[0])}")
    narr = (C.c_int32 * ndim)(*[int(n) for n in ns])
    L.check(fn(ndim, _ptr_array(qs), nq, _ptr_array(knots), narr, periodic_mask, C.byref(params), 0,
               _ptr(order), _ptr(ws), _stream()), "ipx_cell_order")
    return order


def _permute_dev(src: torch.Tensor, order: torch.Tensor, batch: int, gather: bool,
                 dst: torch.Tensor | None = None) -> torch.Tensor:
    lib = L.load()
    n = order.numel()
    if dst is None:
        dst = torch.empty_like(src)
    fn = getattr(lib, f"ipx_{'gather' if gather else 'scatter'}_{_suffix(src)}")
    L.check(fn(_ptr(src), _ptr(order), n, batch, _ptr(dst), _stream()), "ipx_gather/scatter")
    return dst


def _cell_sort_dev(ndim, qs, knots, ns, periodic_mask, params):
    """(order, queries in that order): ipx_cell_sort_*."""
    lib = L.load()
    nq = qs[0].numel()
    order = torch.empty(nq, dtype=torch.int32, device=qs[0].device)  # holds uint32 values
    out = [torch.empty_like(q) for q in qs]
    wbytes = int(lib.ipx_cell_sort_workspace_bytes(ndim, nq, qs[0].element_size()))
    ws = torch.empty(max(wbytes, 256), dtype=torch.uint8, device=qs[0].device)
    fn = getattr(lib, f"ipx_cell_sort_{_suffix(qs[0])}")
    narr = (C.c_int32 * ndim)(*[int(n) for n in ns])
    L.check(fn(ndim, _ptr_array(qs), nq, _ptr_array(knots), narr, periodic_mask, C.byref(params), 0,
               _ptr(order), _ptr_array(out), _ptr(ws), _stream()), "ipx_cell_sort")
    return order, out


def _presorted_stencil(fallback, ndim, knots, sknots, ns, sort_ns, sort_mask, halo, code, c, mask, params):
    """The pre-sorted call of the evaluation from f with its permutation passes folded into the
    kernel: cell keys + radix sort (ipx_cell_order_records_*), then ONE launch that reads query
    order[k] from the interleaved coordinate records and writes out[order[k]]
    (ipx_eval_stencil_perm_*).  Bit-identical to the plain call; `fallback` serves what the stencil
    kernels do not (knot vectors too long for shared memory)."""
    lib = L.load()

    def run(dq, dout, want_idx):
        errorif(want_idx, NotImplementedError, "interval indices are not available with presort")
        nq = dq[0].numel()
        if nq == 0:
            return fallback(dq, dout, False)
        dev, dt = dq[0].device, dq[0].dtype
        order = torch.empty(nq, dtype=torch.int32, device=dev)  # holds uint32 values
        R = 4 if ndim == 3 else ndim
        recs = torch.empty(nq * R, dtype=dt, device=dev)
        ws = torch.empty(max(int(lib.ipx_cell_order_workspace_bytes(nq)), 256), dtype=torch.uint8, device=dev)
        suf = _suffix(dq[0])
        narr = (C.c_int32 * ndim)(*[int(n) for n in sort_ns])
        L.check(getattr(lib, f"ipx_cell_order_records_{suf}")(
            ndim, _ptr_array(dq), nq, _ptr_array(knots), narr, sort_mask, C.byref(params), 0, _ptr(order),
            _ptr(recs), _ptr(ws), _stream()), "ipx_cell_order_records")
        out = dout if dout is not None else torch.empty((nq, 1), dtype=dt, device=dev)
        narr2 = (C.c_int32 * ndim)(*[int(n) for n in ns])
        rc = getattr(lib, f"ipx_eval_stencil_perm_{suf}")(
            ndim, _ptr_array(dq), nq, _ptr_array(knots), _ptr_array(sknots), narr2, _ptr(halo), code, float(c), mask,
            C.byref(params), 0, _ptr(order), _ptr(recs), _ptr(out), _stream())
        if rc == L.IPX_EUNSUPPORTED:
            return fallback(dq, dout, False)
        L.check(rc, "ipx_eval_stencil_perm")
        return out, None

    return run


def _presorted(launch, ndim, knots, ns, mask, params, batch):
    """Wrap `launch` so that it visits the queries in cell order: gather -> evaluate -> scatter
    (results are bit-identical to the plain call; only the memory access pattern changes)."""

    def run(dq, dout, want_idx):
        errorif(want_idx, NotImplementedError, "interval indices are not available with presort")
        nq = dq[0].numel()
        if nq == 0:
            return launch(dq, dout, False)
        order, qs = _cell_sort_dev(ndim, dq, knots, ns, mask, params)
        tmp, _ = launch(qs, None, False)
        if dout is None:
            dout = torch.empty((nq, batch), dtype=tmp.dtype, device=tmp.device)
        _permute_dev(tmp, order, batch, False, dout)
        return dout, None

    return run


def _make_params(ndim, derivative, extrap6, periods) -> L.Params:
    p = L.Params()
    for ax in range(ndim):
        a = p.axis[ax]
        a.derivative = int(derivative[ax])
        a.keep_lo, a.fill_lo = _keep_fill(extrap6[2 * ax])
        a.keep_hi, a.fill_hi = _keep_fill(extrap6[2 * ax + 1])
        a.period = abs(float(periods[ax])) if periods[ax] is not None else 0.0
    return p


def _eval_dev(ndim, qs, knots, perms, ns, tables, layout, batch, method_class, periodic_mask,
              params, return_index=False, out=None):
    """One call into ipx_eval{1,2,3}d_* on the current stream; device tensors only."""
    lib = L.load()
    nq = qs[0].numel()
    dt = qs[0].dtype
    if out is None:
        out = torch.empty((nq, batch), dtype=dt, device=qs[0].device)
    idx = torch.empty((ndim, nq), dtype=torch.int32, device=qs[0].device) if return_index else None
    fn = getattr(lib, f"ipx_eval{ndim}d_{_suffix(qs[0])}")
    args = [_ptr(q) for q in qs] + [nq] + [_ptr(k) for k in knots] + [_ptr(p) for p in perms]
    args += [int(n) for n in ns]
    args += [_ptr_array(tables), layout, batch, method_class, periodic_mask, C.byref(params), 0,
             _ptr(out), _ptr(idx), _stream()]
    L.check(fn(*args), f"ipx_eval{ndim}d")
    return out, idx


# methods whose slope chain is a linear three-point stencil along one axis: evaluated straight
# from f by ipx_eval_stencil_* (no slope tables, no records): name -> slope code
_STENCIL_METHODS = {"cubic": L.IPX_SLOPE_CUBIC, "cardinal": L.IPX_SLOPE_CARDINAL, "catmull-rom": L.IPX_SLOPE_CARDINAL}
_FAST_CALL_MAX = 1 << 16  # queries: larger calls need the device-side choice of kernel (scratch) anyway
_STENCIL_SMEM = 227 * 1024 - 4096  # what the knots + coefficients of all axes may take per CTA
stencil_hint = L.IPX_STENCIL_AUTO  # experiments / benchmarks: force a kernel shape


def _stencil_fits(nks, elem) -> bool:
    """Shared memory the stencil kernels need for the knots and coefficients (st_layout)."""
    total = sum(((nk * elem + 31) & ~31) + nk * 4 * elem for nk in nks)
    return total <= _STENCIL_SMEM


def _halo_dev(ndim: int, f: torch.Tensor, periodic: list, perms: list, pad_first: bool) -> torch.Tensor:
    """f -> halo table (ipx_halo_table_*): two extra rows each side of a periodic axis (wrapped;
    ``pad_first``: the functional form's convention, ghost rows beyond the padded ends), one ghost
    row each side of an open one."""
    lib = L.load()
    ns = [int(f.shape[ax]) for ax in range(ndim)]
    mask = 0
    for ax in range(ndim):
        if periodic[ax]:
            mask |= L.IPX_PAD_FIRST(ax) if pad_first else L.IPX_PERIODIC(ax)
    narr = (C.c_int32 * ndim)(*ns)
    elems = int(lib.ipx_halo_table_elems(ndim, narr, mask, f.element_size()))
    errorif(elems <= 0, RuntimeError, "ipx_halo_table_elems failed")
    out = torch.empty(elems, dtype=f.dtype, device=f.device)
    fn = getattr(lib, f"ipx_halo_table_{_suffix(f)}")
    L.check(fn(ndim, _ptr(f), narr, mask, _ptr_array(list(perms)), _ptr(out), _stream()), "ipx_halo_table")
    return out


def _eval_stencil_dev(ndim, qs, knots, sknots, ns, halo, code, c, mask, params, return_index=False, out=None,
                      records=None):
    """One call into ipx_eval_stencil_* on the current stream; returns None if the C ABI says
    IPX_EUNSUPPORTED (knot vectors too long for shared memory).  ``records``: the packed records
    of the same tables, if the caller has them (the device then picks the faster form per stream)."""
    lib = L.load()
    nq = qs[0].numel()
    dt = qs[0].dtype
    dev = qs[0].device
    if out is None:
        out = torch.empty((nq, 1), dtype=dt, device=dev)
    idx = torch.empty((ndim, nq), dtype=torch.int32, device=dev) if return_index else None
    # (the device-side choice of kernel is made from 2^16 queries up: smaller calls need no scratch)
    ws = torch.empty(4, dtype=torch.int32, device=dev) if stencil_hint == L.IPX_STENCIL_AUTO and nq >= (1 << 16) else None
    fn = getattr(lib, f"ipx_eval_stencil_{_suffix(qs[0])}")
    narr = (C.c_int32 * ndim)(*[int(n) for n in ns])
    rc = fn(ndim, _ptr_array(qs), nq, _ptr_array(knots), _ptr_array(sknots), narr, _ptr(halo), _ptr(records), code,
            float(c), mask, C.byref(params), 0, _ptr(out), _ptr(idx), stencil_hint, _ptr(ws), _stream())
    if rc == L.IPX_EUNSUPPORTED:
        return None
    L.check(rc, "ipx_eval_stencil")
    return out, idx


# --------------------------------------------------------------------------- host <-> device streaming
# Queries that live in host memory (NumPy arrays, CPU tensors) are streamed through the GPU in
# chunks on a few CUDA streams, so the H2D copy of chunk c+1, the kernel of chunk c and the
# D2H copy of chunk c-1 overlap (PCIe is full duplex).  Pinned buffers make the copies truly
# asynchronous; pageable ones still work.
_HOST_CHUNK = 1 << 22            # queries per chunk at most
_HOST_CHUNK_MIN = 1 << 16
_HOST_OUT_BYTES = 256 << 20      # result staging per stream: the chunk shrinks for wide batches
_HOST_STREAMS = 3
_HOST_CACHE_BYTES = 2 << 30      # device memory the cached pipes may hold altogether
_pipes: dict = {}                # (thread, device) -> _HostPipe
_pipes_lock = threading.Lock()


def _host_chunk(batch: int, elem: int) -> int:
    """Queries per chunk: as many as fit _HOST_OUT_BYTES of results, within [2^16, 2^22]."""
    return int(min(_HOST_CHUNK, max(_HOST_CHUNK_MIN, _HOST_OUT_BYTES // max(1, batch * elem))))


def launches_per_host_call(nq: int, batch: int = 1, elem: int = 8) -> int:
    """Chunks (evaluation calls) one host-resident call of nq queries makes."""
    return max(1, -(-int(nq) // _host_chunk(batch, elem)))


class _HostPipe:
    """Streams + flat device staging buffers of one (thread, device); buffers are byte arrays
    viewed per call, so one pipe serves every dtype, dimension and batch width."""

    def __init__(self, dev):
        self.dev = dev
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(_HOST_STREAMS)]
        self.qbuf = [None] * _HOST_STREAMS
        self.obuf = [None] * _HOST_STREAMS

    def nbytes(self):
        return sum(b.numel() for b in self.qbuf + self.obuf if b is not None)

    def views(self, k, tdt, ndim, chunk, batch):
        elem = torch.empty((), dtype=tdt).element_size()
        need_q, need_o = ndim * chunk * elem, chunk * batch * elem
        if self.qbuf[k] is None or self.qbuf[k].numel() < need_q:
            self.qbuf[k] = torch.empty(need_q, dtype=torch.uint8, device=self.dev)
        if self.obuf[k] is None or self.obuf[k].numel() < need_o:
            self.obuf[k] = torch.empty(need_o, dtype=torch.uint8, device=self.dev)
        q = self.qbuf[k][:need_q].view(tdt).view(ndim, chunk)
        o = self.obuf[k][:need_o].view(tdt)
        return q, o


def _host_pipe(dev) -> _HostPipe:
    key = (threading.get_ident(), dev.index)
    with _pipes_lock:
        pipe = _pipes.get(key)
        if pipe is None:
            # bound what the cache holds: drop other threads' / devices' pipes first
            while _pipes and sum(p.nbytes() for p in _pipes.values()) > _HOST_CACHE_BYTES:
                _pipes.pop(next(iter(_pipes)))
            pipe = _pipes[key] = _HostPipe(dev)
        return pipe


def _run_host(dev, qd, batch, launch, out_host):
    """qd: 1-D CPU tensors; out_host: (nq, batch) CPU tensor; launch(dq, dout) enqueues the
    kernel on the current stream."""
    ndim = len(qd)
    nq = qd[0].numel()
    if nq == 0:
        return
    pipe = _host_pipe(dev)
    tdt = qd[0].dtype
    chunk = _host_chunk(batch, qd[0].element_size())
    cur = torch.cuda.current_stream()
    for s in pipe.streams:
        s.wait_stream(cur)
    for c, lo in enumerate(range(0, nq, chunk)):
        hi = min(nq, lo + chunk)
        m = hi - lo
        k = c % _HOST_STREAMS
        with torch.cuda.stream(pipe.streams[k]):
            qv, ov = pipe.views(k, tdt, ndim, chunk, batch)
            dq = [qv[ax][:m] for ax in range(ndim)]
            for ax in range(ndim):
                dq[ax].copy_(qd[ax][lo:hi], non_blocking=True)
            dout = ov[: m * batch].view(m, batch)
            launch(dq, dout)
            out_host[lo:hi].copy_(dout, non_blocking=True)
    for s in pipe.streams:
        s.synchronize()
    if pipe.nbytes() > _HOST_CACHE_BYTES:  # an unusually wide call: do not keep its buffers
        with _pipes_lock:
            _pipes.pop((threading.get_ident(), dev.index), None)


# --------------------------------------------------------------------------- front end
class _Inputs:
    """dtype / shape front matter shared by the functional and class forms
    (_spline.py:503-521, 663-689, 882-910)."""

    def __init__(self, qs, knots, f_dtype_item, stored=None):
        self.dev = _device()
        if stored is not None and all(isinstance(q, torch.Tensor) and q.dtype == stored.tdt for q in qs):
            # queries already in the stored tables' type: nothing to promote (the common call of
            # Interpolator*.__call__ with device tensors; saves the NumPy dtype algebra per call)
            self.want_torch = True
            self.xdt, self.fdt, self.rdt, self.tdt = stored.xdt, stored.fdt, stored.rdt, stored.tdt
            self.complex = stored.complex
            if np.result_type(stored.xdt, stored.rdt) == stored.xdt:
                return
        self.want_torch = any(isinstance(a, torch.Tensor) for a in (*qs, *knots)) or (
            stored.want_torch if stored is not None else isinstance(f_dtype_item, torch.Tensor))
        self.xdt = _promote([*knots, *qs])
        errorif(np.issubdtype(self.xdt, np.complexfloating), TypeError, "knots and queries must be real")
        fdt = stored.fdt if stored is not None else _promote([f_dtype_item])
        self.fdt = np.dtype(np.result_type(fdt, self.xdt))
        self.complex = bool(np.issubdtype(self.fdt, np.complexfloating))
        # the kernels compute in one real type T = the real part of result_type(f, x)
        self.rdt = np.dtype(np.float64) if self.fdt in (np.float64, np.complex128) else np.dtype(np.float32)
        self.tdt = _NP2T[self.rdt]

    def coords(self, a):
        t = _to_dev(a, self.xdt, self.dev)
        return t if t.dtype == self.tdt else t.to(self.tdt)

    def table(self, a):
        t = _to_dev(a, self.fdt, self.dev)
        if self.complex:
            t = torch.view_as_real(t).contiguous()
        return t if t.dtype == self.tdt else t.to(self.tdt)

    def queries(self, qs):
        """-> (list of flat 1-D tensors, query shape, host_mode).  Host-resident queries stay on
        the host (they are streamed); if any query is a CUDA tensor all go to the device."""
        ts = []
        for q in qs:
            if isinstance(q, torch.Tensor):
                t = q
            else:
                t = torch.from_numpy(np.array(q, dtype=self.xdt, order="C"))
            ts.append(t)
        host = all(not t.is_cuda for t in ts)
        if not host:
            ts = [t.to(self.dev) for t in ts]
        ts = [t if t.dtype == self.tdt else t.to(self.tdt) for t in ts]
        if len(ts) > 1:
            ts = list(torch.broadcast_tensors(*ts))
        shape = tuple(ts[0].shape)
        ts = [t.contiguous().reshape(-1) for t in ts]
        return ts, shape, host

    def finish(self, out: torch.Tensor, outshape, host):
        if self.complex:
            out = torch.view_as_complex(out.reshape(tuple(outshape) + (2,)))
        out = out.reshape(tuple(outshape))
        tdt = _NP2T[self.fdt]
        if out.dtype != tdt:
            out = out.to(tdt)
        if self.want_torch:
            return out
        return (out if host else out.cpu()).numpy()


def _execute(io, qs, launch, batch, tail_shape, out=None, return_index=False):
    """Run `launch` over the queries, on the device or streamed from the host."""
    qd, qshape, host = io.queries(qs)
    outshape = qshape + tuple(tail_shape)
    nq = qd[0].numel()
    if return_index and host:
        qd = [q.to(io.dev) for q in qd]
        host = False
    if out is not None:
        errorif(not isinstance(out, torch.Tensor) or out.dtype != io.tdt or out.numel() != nq * batch
                or not out.is_contiguous() or out.is_cuda == host, ValueError,
                "out must be a contiguous tensor of the result dtype and size, on the side the queries are on")
        out = out.view(nq, batch)
    if host:
        if out is None:
            out = torch.empty((nq, batch), dtype=io.tdt)
        _run_host(io.dev, qd, batch, lambda dq, dout: launch(dq, dout, False), out)
        return io.finish(out, outshape, True)
    res, idx = launch(qd, out, return_index)
    final = io.finish(res, outshape, False)
    if return_index:
        return final, (idx if io.want_torch else idx.cpu().numpy())
    return final


def _shape(a):
    return tuple(a.shape) if hasattr(a, "shape") else np.shape(a)


def _check_shapes(ndim, knots, fshape):
    """The reference's argument checks (_spline.py:516-521, 679-689, 887-902); they need no
    device, so bad calls fail the same way with or without a GPU."""
    errorif(len(fshape) < ndim, ValueError, "x and f must be arrays of equal length")
    for ax, name in zip(range(ndim), "xyz"):
        ks = _shape(knots[ax])
        errorif(
            (len(ks) != 1) or (ks[0] != fshape[ax]), ValueError,
            f"{name} and f must be arrays of equal length",
        )


def _method_class(method):
    return L.IPX_NEAREST if method == "nearest" else L.IPX_LINEAR if method == "linear" else L.IPX_CUBIC


def _interp(ndim, qs, knots, f, method, derivative, extrap, period, kwargs, return_index=False):
    """Functional form.  Body of interp1d/2d/3d (_spline.py:503-592, 663-805, 882-1105)."""
    methods = (METHODS_1D, METHODS_2D, METHODS_3D)[ndim - 1]
    names = TABLES[ndim]
    kwargs = dict(kwargs)
    given = {n: kwargs.pop(n, None) for n in names[1:]}
    out_arg = kwargs.pop("out", None)
    presort = bool(kwargs.pop("presort", False))
    if ndim == 1:
        axis = kwargs.get("axis", 0)
        errorif(axis != 0, NotImplementedError, "only axis=0 is supported")
    kwargs.pop("axis", None)

    fshape_user = _shape(f)
    _check_shapes(ndim, knots, fshape_user)
    errorif(method not in methods, ValueError, f"unknown method {method}")
    io = _Inputs(qs, knots, f)
    kn = [io.coords(k) for k in knots]

    periods = _parse_ndarg(period, ndim)
    derivs = _parse_ndarg(derivative, ndim)
    ex = _parse_extrap(extrap, ndim)

    tabs = {"f": io.table(f)}
    for n in names[1:]:
        tabs[n] = None if given[n] is None else io.table(given[n])
    batch = int(np.prod(tabs["f"].shape[ndim:], dtype=np.int64))
    mclass = _method_class(method)

    pk = [None] * ndim  # per periodic axis: (xpad, perm)
    for ax in range(ndim):
        if periods[ax] is not None:
            pk[ax] = _periodic_knots_dev(kn[ax], periods[ax])

    mask = 0
    perms = [None] * ndim
    kuse = list(kn)
    packed = None
    params = _make_params(ndim, derivs, ex, periods)
    if (method in _STENCIL_METHODS and batch == 1 and not io.complex
            and all(tabs[n] is None for n in names[1:])
            and _stencil_fits([kn[ax].shape[0] + (2 if pk[ax] is not None else 0) for ax in range(ndim)],
                              kn[0].element_size())):
        # linear three-point stencils: evaluate straight from f.  The functional form takes its
        # slopes on the PADDED knots (_spline.py:924-938 precede :1027-1040): IPX_PAD_FIRST.
        per = [p is not None for p in pk]
        halo = _halo_dev(ndim, tabs["f"], per, [p[1] if p is not None else None for p in pk], True)
        sk = [pk[ax][0] if per[ax] else kn[ax] for ax in range(ndim)]
        smask = sum(L.IPX_PAD_FIRST(ax) for ax in range(ndim) if per[ax])
        sns = [tabs["f"].shape[ax] for ax in range(ndim)]
        code = _STENCIL_METHODS[method]
        cten = float(kwargs.get("c", 0)) if method == "cardinal" else 0.0
        # 2-D, several queries per cell: the records of the (padded) table as well, one fused pass; the
        # device then picks the faster form for the stream it sees (random order).  In 3-D the
        # records are not worth building per call: dense cell-sorted streams run the column shape
        # of the evaluation from f (2.2 ms per 1e8 queries on 256^3 against 1.9 ms on records that
        # take 0.5 ms to build)
        recs = None
        nq_total = int(np.prod([int(v) for v in np.broadcast_shapes(*[_shape(q) for q in qs])], dtype=np.int64))
        cells = float(np.prod([max(1, int(k.shape[0]) - 1) for k in sk], dtype=np.float64))
        if ndim == 2 and nq_total >= 3.0 * cells and nq_total >= (1 << 16):
            ftab = tabs["f"]
            for ax in range(ndim):
                if per[ax]:
                    ftab = _pad_periodic_dev(ftab, ax, pk[ax][1])
            recs = _slopes_pack_dev(ndim, sk, ftab, 1, method, kwargs)

        def launch_st(dq, dout, want_idx):
            res = _eval_stencil_dev(ndim, dq, sk, [None] * ndim, sns, halo, code, cten, smask, params,
                                    want_idx, dout, recs)
            errorif(res is None, RuntimeError, "ipx_eval_stencil: unsupported")
            return res

        if presort:
            # the cell sort sees the padded knots as they are: wrap-only axes of padded extent
            pns = [n + (2 if per[ax] else 0) for ax, n in enumerate(sns)]
            pmask = sum(L.IPX_WRAP_ONLY(ax) for ax in range(ndim) if per[ax])
            launch_st = _presorted_stencil(_presorted(launch_st, ndim, sk, pns, pmask, params, batch), ndim, sk,
                                           [None] * ndim, sns, pns, pmask, halo, code, cten, smask, params)
        return _execute(io, qs, launch_st, batch, fshape_user[ndim:], out_arg, return_index)
    if mclass == L.IPX_CUBIC:
        missing = [n for n in names[1:] if tabs[n] is None]
        if missing and any(p is not None for p in pk):
            # the functional form differentiates the PADDED tables (_spline.py:924-938 then
            # :1027-1040): materialise the pad, then evaluate with wrap-only axes.
            for ax in range(ndim):
                if pk[ax] is None:
                    continue
                for n in names:
                    if tabs[n] is not None:
                        tabs[n] = _pad_periodic_dev(tabs[n], ax, pk[ax][1])
                kuse[ax] = pk[ax][0]
                mask |= L.IPX_WRAP_ONLY(ax)
        else:
            for ax in range(ndim):
                if pk[ax] is not None:
                    kuse[ax], perms[ax] = pk[ax]
                    mask |= L.IPX_PERIODIC(ax)
        if len(missing) == len(names) - 1:
            # no slope table supplied: f (padded where periodic) -> records in one launch
            packed = _slopes_pack_dev(ndim, kuse, tabs["f"], batch, method, kwargs)
        if packed is None:
            for name, src, ax in CHAIN[ndim]:
                if tabs[name] is None:
                    tabs[name] = _slopes_dev(kuse[ax], tabs[src], method, ax, kwargs, io.complex)
            for n in names[1:]:
                errorif(tabs[n].shape != tabs["f"].shape, ValueError,
                        f"{n} must have the same shape as f")
            tables = [tabs[n] for n in names]
    else:
        for ax in range(ndim):
            if pk[ax] is not None:
                kuse[ax], perms[ax] = pk[ax]
                mask |= L.IPX_PERIODIC(ax)
        tables = [tabs["f"]]

    ns = [tabs["f"].shape[ax] for ax in range(ndim)]

    # Many queries on scalar-valued cubic tables: interleave the tables once (one pass over
    # them) and use the tile kernel, as Interpolator* does; few queries: read the reference's
    # separate tables in place.
    layout = L.IPX_SOA
    nq_total = int(np.prod([int(v) for v in np.broadcast_shapes(*[_shape(q) for q in qs])], dtype=np.int64))
    if mclass == L.IPX_CUBIC and packed is not None:
        tables, layout = [packed], L.IPX_AOS
    elif mclass == L.IPX_CUBIC and batch == 1 and nq_total >= (1 << 16) and 2 * nq_total >= tables[0].numel():
        tables, layout = [_pack_dev(ndim, tables)], L.IPX_AOS

    def launch(dq, dout, want_idx):
        return _eval_dev(ndim, dq, kuse, perms, ns, tables, layout, batch, mclass, mask, params,
                         want_idx, dout)

    if presort:
        launch = _presorted(launch, ndim, kuse, ns, mask, params, batch)
    return _execute(io, qs, launch, batch, fshape_user[ndim:], out_arg, return_index)


def interp1d(xq, x, f, method="cubic", derivative=0, extrap=False, period=None, **kwargs):
    """Interpolate a 1d function.  Same contract as interpax.interp1d (_spline.py:445-592):
    ``xq`` (Nq,), ``x`` (Nx,), ``f`` (Nx, ...), ``method`` one of nearest / linear / cubic /
    cubic2 / catmull-rom / cardinal / monotonic / monotonic-0 / akima, ``derivative`` >= 0,
    ``extrap`` bool | float | (lo, hi), ``period`` None | float, kwargs ``fx``, ``c``,
    ``bc_type``.  Returns (Nq, ...)."""
    return _interp(1, (xq,), (x,), f, method, derivative, extrap, period, kwargs)


def interp2d(xq, yq, x, y, f, method="cubic", derivative=0, extrap=False, period=None, **kwargs):
    """Interpolate a 2d function.  Same contract as interpax.interp2d (_spline.py:596-805);
    kwargs ``fx``, ``fy``, ``fxy``, ``c``, ``bc_type``."""
    return _interp(2, (xq, yq), (x, y), f, method, derivative, extrap, period, kwargs)


def interp3d(xq, yq, zq, x, y, z, f, method="cubic", derivative=0, extrap=False, period=None,
             **kwargs):
    """Interpolate a 3d function.  Same contract as interpax.interp3d (_spline.py:809-1105);
    kwargs ``fx`` ... ``fxyz``, ``c``, ``bc_type``."""
    return _interp(3, (xq, yq, zq), (x, y, z), f, method, derivative, extrap, period, kwargs)


def approx_df(x, f, method="cubic", axis=-1, **kwargs):
    """First derivatives at the knots.  Same contract as interpax.approx_df
    (_fd_derivs.py:9-76)."""
    want_torch = isinstance(x, torch.Tensor) or isinstance(f, torch.Tensor)
    errorif(method not in _SLOPE_CODE, ValueError, f"got unknown method {method}")
    dev = _device()
    xdt = _promote([x])
    # dxi * df promotes with the knots (_fd_derivs.py:84-90): float32 data on float64 knots gives
    # a float64 result, and the knots are never rounded down before they are differenced
    fdt = np.dtype(np.result_type(_promote([f]), xdt))
    cplx = bool(np.issubdtype(fdt, np.complexfloating))
    rdt = np.dtype(np.float64) if fdt in (np.float64, np.complex128) else np.dtype(np.float32)
    ft = _to_dev(f, fdt, dev)
    nd = ft.ndim
    axis = axis % nd
    if cplx:
        ft = torch.view_as_real(ft).contiguous()
    xt = _to_dev(x, xdt, dev).to(_NP2T[rdt])
    out = _slopes_dev(xt, ft, method, axis, kwargs, cplx)
    if cplx:
        out = torch.view_as_complex(out)
    return out if want_torch else out.cpu().numpy()


def vjp_tables(qs, knots, fshape, cotangent, method="cubic", derivative=0, extrap=False, period=None):
    """Reverse-mode rule of ``interp{1,2,3}d`` with respect to its tables.

    ``qs`` / ``knots``: tuples of the query and knot arrays (1, 2 or 3 of each); ``fshape`` the
    shape of ``f``; ``cotangent`` an array of the output's shape.  Returns a dict of gradients
    with the shape of ``f``: ``{"f": ...}`` for ``method="linear"`` and ``{"f", "fx", ...}`` (the
    reference's table names) for the cubic family, i.e. the transpose of the map
    tables -> output that JAX derives by differentiating the gathers and the ``A @ F``
    contraction (interpax/_spline.py:581-589, 779-800, 1070-1099).  The dependence of slope
    tables on ``f`` (``approx_df``) is not included: this is the rule for the tables as inputs.
    """
    ndim = len(knots)
    errorif(ndim not in (1, 2, 3) or len(qs) != ndim, ValueError, "need 1, 2 or 3 query and knot arrays")
    errorif(method == "nearest", NotImplementedError, "nearest-neighbour selection has no table gradient")
    errorif(method not in METHODS_1D, ValueError, f"unknown method {method}")
    fshape = tuple(int(v) for v in fshape)
    _check_shapes(ndim, knots, fshape)
    io = _Inputs(qs, knots, cotangent)
    errorif(io.complex, NotImplementedError, "complex cotangents: pass real and imaginary parts separately")
    kn = [io.coords(k) for k in knots]
    qd, qshape, _ = io.queries(qs)
    qd = [q.to(io.dev) for q in qd]
    batch = int(np.prod(fshape[ndim:], dtype=np.int64))
    cot = _to_dev(cotangent, io.rdt, io.dev).reshape(-1)
    errorif(cot.numel() != qd[0].numel() * batch, ValueError, "cotangent must have the shape of the output")
    periods = _parse_ndarg(period, ndim)
    derivs = _parse_ndarg(derivative, ndim)
    ex = _parse_extrap(extrap, ndim)
    mask = 0
    perms = [None] * ndim
    kuse = list(kn)
    for ax in range(ndim):
        if periods[ax] is not None:
            kuse[ax], perms[ax] = _periodic_knots_dev(kn[ax], periods[ax])
            mask |= L.IPX_PERIODIC(ax)
    mclass = _method_class(method)
    names = TABLES[ndim] if mclass == L.IPX_CUBIC else ("f",)
    grads = [torch.zeros(fshape, dtype=io.tdt, device=io.dev) for _ in names]
    params = _make_params(ndim, derivs, ex, periods)
    lib = L.load()
    fn = getattr(lib, f"ipx_vjp_tables_{_suffix(qd[0])}")
    narr = (C.c_int32 * ndim)(*[int(fshape[ax]) for ax in range(ndim)])
    L.check(fn(ndim, _ptr_array(qd), qd[0].numel(), _ptr_array(kuse), _ptr_array(perms), narr, batch, mclass,
               mask, C.byref(params), 0, _ptr(cot), _ptr_array(grads), _stream()), "ipx_vjp_tables")
    out = dict(zip(names, grads))
    return out if io.want_torch else {k: v.cpu().numpy() for k, v in out.items()}


def _slope_opts(method, kwargs):
    """ipx_slope_opts of a reverse-mode call: tension and boundary KINDS (the boundary values are
    constants of the map f -> slopes and drop out of its transpose)."""
    opts = L.SlopeOpts()
    opts.c = float(kwargs.get("c", 0)) if method == "cardinal" else 0.0
    if method == "cubic2":
        bc = kwargs.get("bc_type", "not-a-knot")
        kinds = []
        for b in ((bc, bc) if isinstance(bc, str) else tuple(bc)):
            if isinstance(b, str):
                errorif(b not in ("clamped", "natural", "not-a-knot"), ValueError, f"bc_type={b} is not allowed.")
                kinds.append({"clamped": L.IPX_BC_FIRST, "natural": L.IPX_BC_SECOND,
                              "not-a-knot": L.IPX_BC_NOT_A_KNOT}[b])
            else:
                errorif(b[0] not in (1, 2), ValueError, "The specified derivative order must be 1 or 2.")
                kinds.append(L.IPX_BC_FIRST if b[0] == 1 else L.IPX_BC_SECOND)
        errorif(len(kinds) != 2, ValueError, "`bc_type` must contain 2 elements to specify start and end conditions.")
        opts.bc_start, opts.bc_end = kinds
    return opts


def _slopes_vjp_dev(x: torch.Tensor, g: torch.Tensor, method: str, axis: int, kwargs: dict,
                    into: torch.Tensor | None = None) -> torch.Tensor:
    """(d approx_df / d f)^T g on device tensors; adds into ``into`` if given."""
    errorif(method not in ("cubic", "cubic2", "cardinal", "catmull-rom"), NotImplementedError,
            "approx_df is linear in f for method = cubic, cubic2, cardinal, catmull-rom only; "
            "monotonic and akima have no transpose here")
    lib = L.load()
    axis = axis % g.ndim
    n = g.shape[axis]
    errorif(x.ndim != 1 or x.shape[0] != n, ValueError, "x and f must be arrays of equal length")
    outer = int(np.prod(g.shape[:axis], dtype=np.int64))
    inner = int(np.prod(g.shape[axis + 1:], dtype=np.int64))
    code = _SLOPE_CODE[method]
    opts = _slope_opts(method, kwargs)
    g = g.contiguous()
    out = torch.empty_like(g) if into is None else into
    wbytes = lib.ipx_slopes_vjp_workspace_bytes(code, n, outer, inner, g.element_size())
    ws = torch.empty(max(int(wbytes), 16), dtype=torch.uint8, device=g.device)
    fn = getattr(lib, f"ipx_slopes_vjp_{_suffix(g)}")
    L.check(fn(_ptr(x), n, _ptr(g), outer, inner, code, C.byref(opts), 0 if into is None else 1, _ptr(out),
               _ptr(ws), _stream()), "ipx_slopes_vjp")
    return out


def approx_df_vjp(x, cotangent, method="cubic", axis=-1, **kwargs):
    """Reverse-mode rule of ``approx_df(x, f, method, axis)`` with respect to ``f``: the transpose of
    the (linear) slope map applied to ``cotangent`` (an array of the shape of ``f``).  method = cubic,
    cubic2 (``bc_type`` as for ``approx_df``; boundary values are constants), cardinal, catmull-rom.
    What ``jax.vjp(lambda f: approx_df(x, f, method, axis), f)[1]`` returns in the reference
    (_fd_derivs.py:79-101, 160-300, 303-324)."""
    want_torch = isinstance(x, torch.Tensor) or isinstance(cotangent, torch.Tensor)
    dev = _device()
    xdt = _promote([x])
    gdt = np.dtype(np.result_type(_promote([cotangent]), xdt))
    errorif(np.issubdtype(gdt, np.complexfloating), NotImplementedError,
            "complex cotangents: pass real and imaginary parts separately")
    rdt = np.dtype(np.float64) if gdt == np.float64 else np.dtype(np.float32)
    gt = _to_dev(cotangent, rdt, dev)
    xt = _to_dev(x, xdt, dev).to(_NP2T[rdt])
    out = _slopes_vjp_dev(xt, gt, method, axis % gt.ndim, kwargs)
    return out if want_torch else out.cpu().numpy()


def vjp_f(qs, knots, fshape, cotangent, method="cubic", derivative=0, extrap=False, period=None,
          functional=True, **kwargs):
    """Reverse-mode rule of ``interp{1,2,3}d(q, knots, f, method)`` (``functional=True``) or
    ``Interpolator{1,2,3}D(knots, f, method)(q)`` (``functional=False``) with respect to ``f``, the
    slope chain included, for every method that is linear in ``f``: linear, cubic, cubic2 (any
    ``bc_type``), cardinal, catmull-rom.  The transpose of the evaluation with respect to its 2^d
    tables (``vjp_tables``) is followed by the transposed links of the reference's slope chain in
    reverse order (_spline.py:1027-1040: fxyz <- fxy <- fx <- f ...), each a transposed stencil or,
    for cubic2, a transposed tridiagonal solve.  On periodic axes the functional form takes its
    slopes on the padded tables (_spline.py:924-938): the chain runs on the padded shape and the pad
    is transposed last.  Returns the gradient, an array of shape ``fshape``."""
    ndim = len(knots)
    errorif(ndim not in (1, 2, 3) or len(qs) != ndim, ValueError, "need 1, 2 or 3 query and knot arrays")
    errorif(method not in ("linear", "cubic", "cubic2", "cardinal", "catmull-rom"), NotImplementedError,
            "the evaluation is linear in f for method = linear, cubic, cubic2, cardinal, catmull-rom; "
            "nearest, monotonic and akima are not covered")
    fshape = tuple(int(v) for v in fshape)
    _check_shapes(ndim, knots, fshape)
    io = _Inputs(qs, knots, cotangent)
    errorif(io.complex, NotImplementedError, "complex cotangents: pass real and imaginary parts separately")
    kn = [io.coords(k) for k in knots]
    qd, qshape, _ = io.queries(qs)
    qd = [q.to(io.dev) for q in qd]
    batch = int(np.prod(fshape[ndim:], dtype=np.int64))
    cot = _to_dev(cotangent, io.rdt, io.dev).reshape(-1)
    errorif(cot.numel() != qd[0].numel() * batch, ValueError, "cotangent must have the shape of the output")
    periods = _parse_ndarg(period, ndim)
    params = _make_params(ndim, _parse_ndarg(derivative, ndim), _parse_extrap(extrap, ndim), periods)
    mclass = _method_class(method)
    names = TABLES[ndim] if mclass == L.IPX_CUBIC else ("f",)
    pad_first = functional and mclass == L.IPX_CUBIC
    mask, perms, kuse, pshape = 0, [None] * ndim, list(kn), list(fshape)
    pk = [None] * ndim
    for ax in range(ndim):
        if periods[ax] is None:
            continue
        pk[ax] = _periodic_knots_dev(kn[ax], periods[ax])
        kuse[ax] = pk[ax][0]
        if pad_first:
            mask |= L.IPX_WRAP_ONLY(ax)
            pshape[ax] += 2
        else:
            perms[ax] = pk[ax][1]
            mask |= L.IPX_PERIODIC(ax)
    grads = [torch.zeros(pshape, dtype=io.tdt, device=io.dev) for _ in names]
    narr = (C.c_int32 * ndim)(*[int(pshape[ax]) for ax in range(ndim)])
    fn = getattr(L.load(), f"ipx_vjp_tables_{_suffix(qd[0])}")
    L.check(fn(ndim, _ptr_array(qd), qd[0].numel(), _ptr_array(kuse), _ptr_array(perms), narr, batch, mclass,
               mask, C.byref(params), 0, _ptr(cot), _ptr_array(grads), _stream()), "ipx_vjp_tables")
    g = dict(zip(names, grads))
    if mclass == L.IPX_CUBIC:
        # slopes are taken on the padded knots (functional form) or on the caller's (classes)
        sk = [kuse[ax] if (pad_first and pk[ax] is not None) else kn[ax] for ax in range(ndim)]
        for name, src, ax in reversed(CHAIN[ndim]):
            _slopes_vjp_dev(sk[ax], g[name], method, ax, kwargs, into=g[src])
    gf = g["f"]
    if pad_first:
        # transpose of the wrap pad (_spline.py:1123-1135): padded row u holds f[perm[(u-1) mod n]]
        for ax in range(ndim):
            if pk[ax] is None:
                continue
            n = fshape[ax]
            perm = pk[ax][1].to(torch.int64)
            src = perm[(torch.arange(n + 2, device=io.dev) - 1) % n]
            shape = list(gf.shape)
            shape[ax] = n
            gf = torch.zeros(shape, dtype=gf.dtype, device=gf.device).index_add_(ax, src, gf)
    return gf if io.want_torch else gf.cpu().numpy()


def _grad_stencil_setup(qs, knots, f, method, derivative, extrap, period, functional, kwargs, dtype_item):
    ndim = len(knots)
    errorif(ndim not in (1, 2, 3) or len(qs) != ndim, ValueError, "need 1, 2 or 3 query and knot arrays")
    errorif(method not in _STENCIL_METHODS, NotImplementedError,
            "knot / f derivatives are available for method = cubic, cardinal, catmull-rom "
            "(local linear stencils); cubic2, monotonic and akima are not covered")
    fshape = _shape(f)
    _check_shapes(ndim, knots, fshape)
    errorif(len(fshape) != ndim, NotImplementedError, "scalar-valued f only")
    io = _Inputs(qs, knots, dtype_item)
    errorif(io.complex, NotImplementedError, "real data only")
    kn = [io.coords(k) for k in knots]
    qd, qshape, _ = io.queries(qs)
    qd = [q.to(io.dev) for q in qd]
    ft = io.table(f)
    periods = _parse_ndarg(period, ndim)
    params = _make_params(ndim, _parse_ndarg(derivative, ndim), _parse_extrap(extrap, ndim), periods)
    mask, kuse, perms = 0, list(kn), [None] * ndim
    for ax in range(ndim):
        if periods[ax] is None:
            continue
        xpad, perm = _periodic_knots_dev(kn[ax], periods[ax])
        kuse[ax] = xpad
        if functional:
            mask |= L.IPX_PAD_FIRST(ax)
            perms[ax] = perm
        else:
            ident = torch.arange(perm.numel(), dtype=perm.dtype, device=perm.device)
            errorif(not bool(torch.equal(perm, ident)), NotImplementedError,
                    "class-form periodic axes need knots sorted inside one period")
            mask |= L.IPX_PERIODIC(ax)
    code = _STENCIL_METHODS[method]
    c = float(kwargs.get("c", 0)) if method == "cardinal" else 0.0
    narr = (C.c_int32 * ndim)(*[int(fshape[ax]) for ax in range(ndim)])
    return io, ndim, qd, qshape, kn, kuse, perms, ft, params, mask, code, c, narr


_CHAIN_KNOT_METHODS = ("cubic2", "monotonic", "monotonic-0")


def _link_vjp(x, src, slopes, g, method, axis, kwargs, g_src, g_x):
    """Reverse rule of one link ``slopes = approx_df(x, src, method, axis)``: adds (d slopes/d src)^T g
    into ``g_src`` and (d slopes/d x)^T g into ``g_x``."""
    lib = L.load()
    n = src.shape[axis]
    outer = int(np.prod(src.shape[:axis], dtype=np.int64))
    inner = int(np.prod(src.shape[axis + 1:], dtype=np.int64))
    code = _SLOPE_CODE[method]
    opts, keep = _slope_opts_values(src, method, axis, kwargs)
    suf = _suffix(src)
    g = g.contiguous()
    y, fwd = g, None
    if method in ("monotonic", "monotonic-0"):
        gf_ptr = _ptr(g_src)  # not linear in f: the knot kernel differentiates with respect to f as well
    else:
        wbytes = lib.ipx_slopes_vjp_workspace_bytes(code, n, outer, inner, g.element_size())
        ws = torch.empty(max(int(wbytes), 16), dtype=torch.uint8, device=g.device)
        L.check(getattr(lib, f"ipx_slopes_vjp_{suf}")(_ptr(x), n, _ptr(g), outer, inner, code, C.byref(opts), 1,
                                                      _ptr(g_src), _ptr(ws), _stream()), "ipx_slopes_vjp")
        gf_ptr = None
        if method == "cubic2":
            # T^-T g, left in the workspace behind the 3 n factor entries
            es = g.element_size()
            y = ws[3 * n * es: (3 * n + g.numel()) * es].view(g.dtype)
            fwd = slopes
    L.check(getattr(lib, f"ipx_slopes_knots_vjp_{suf}")(_ptr(x), n, _ptr(src), _ptr(fwd), _ptr(y), outer, inner,
                                                        code, C.byref(opts), _ptr(g_x), gf_ptr, _stream()),
            "ipx_slopes_knots_vjp")


def _vjp_knots_chain(qs, knots, f, cotangent, method, derivative, extrap, period, functional, kwargs):
    """vjp_knots for the methods evaluated from tables: the evaluation's own knot dependence
    (ipx_vjp_knots_tables_*), the table cotangents (ipx_vjp_tables_*) and every link of the slope
    chain in reverse (ipx_slopes_vjp_* / ipx_slopes_knots_vjp_*)."""
    io, ndim, fshape, kn, qd, qshape, params, pk, mask, perms, kuse, sk, tabs = _chain_setup(
        qs, knots, f, cotangent, method, derivative, extrap, period, functional, kwargs)
    ft = tabs["f"]
    cot = _to_dev(cotangent, io.rdt, io.dev).reshape(-1)
    errorif(cot.numel() != qd[0].numel(), ValueError, "cotangent must have the shape of the output")
    names = TABLES[ndim]
    tables = [tabs[n] for n in names]
    pshape = list(tabs["f"].shape)
    narr = (C.c_int32 * ndim)(*[int(pshape[ax]) for ax in range(ndim)])
    lib = L.load()
    suf = _suffix(ft)
    # table cotangents
    g = {n: torch.zeros_like(tabs["f"]) for n in names}
    L.check(getattr(lib, f"ipx_vjp_tables_{suf}")(ndim, _ptr_array(qd), qd[0].numel(), _ptr_array(kuse),
                                                  _ptr_array(perms), narr, 1, L.IPX_CUBIC, mask, C.byref(params), 0,
                                                  _ptr(cot), _ptr_array([g[n] for n in names]), _stream()),
            "ipx_vjp_tables")
    # the evaluation's own dependence on the (lookup) knots
    g_look = [torch.zeros_like(k) for k in kuse]
    L.check(getattr(lib, f"ipx_vjp_knots_tables_{suf}")(ndim, _ptr_array(qd), qd[0].numel(), _ptr_array(kuse),
                                                        _ptr_array(perms), narr, _ptr_array(tables), mask,
                                                        C.byref(params), 0, _ptr(cot), _ptr_array(g_look), _stream()),
            "ipx_vjp_knots_tables")
    # the chain, last link first
    g_sk = [torch.zeros_like(k) for k in sk]
    for name, src, ax in reversed(CHAIN[ndim]):
        _link_vjp(sk[ax], tabs[src], tabs[name], g[name], method, ax, kwargs, g[src], g_sk[ax])
    gf = g["f"]
    gk = []
    for ax in range(ndim):
        if pk[ax] is None:
            gk.append(g_look[ax] + g_sk[ax])
            continue
        # padded knot u is the caller's knot perm[(u-1) mod n], shifted by a multiple of the period
        n = fshape[ax]
        src_ix = pk[ax][1].to(torch.int64)[(torch.arange(n + 2, device=io.dev) - 1) % n]
        if functional:
            gx = torch.zeros_like(kn[ax]).index_add_(0, src_ix, g_look[ax] + g_sk[ax])
            shape = list(gf.shape)
            shape[ax] = n
            gf = torch.zeros(shape, dtype=gf.dtype, device=gf.device).index_add_(ax, src_ix, gf)
        else:
            gx = torch.zeros_like(kn[ax]).index_add_(0, src_ix, g_look[ax]) + g_sk[ax]
        gk.append(gx)
    out = dict(zip("xyz", gk))
    out["f"] = gf
    return out if io.want_torch else {k: v.cpu().numpy() for k, v in out.items()}


def _chain_setup(qs, knots, f, dtype_item, method, derivative, extrap, period, functional, kwargs):
    """Forward pass shared by the chained reverse and forward rules: tables of the slope chain."""
    ndim = len(knots)
    errorif(ndim not in (1, 2, 3) or len(qs) != ndim, ValueError, "need 1, 2 or 3 query and knot arrays")
    methods = (METHODS_1D, METHODS_2D, METHODS_3D)[ndim - 1]
    errorif(method not in methods, ValueError, f"unknown method {method}")
    fshape = _shape(f)
    _check_shapes(ndim, knots, fshape)
    errorif(len(fshape) != ndim, NotImplementedError, "scalar-valued f only")
    io = _Inputs(qs, knots, dtype_item)
    errorif(io.complex, NotImplementedError, "real data only")
    kn = [io.coords(k) for k in knots]
    qd, qshape, _ = io.queries(qs)
    qd = [q.to(io.dev) for q in qd]
    ft = _to_dev(f, io.rdt, io.dev).contiguous()
    periods = _parse_ndarg(period, ndim)
    ex = _parse_extrap(extrap, ndim)
    params = _make_params(ndim, _parse_ndarg(derivative, ndim), ex, periods)
    pk = [None] * ndim
    mask, perms, kuse, sk = 0, [None] * ndim, list(kn), list(kn)
    tabs = {"f": ft}
    for ax in range(ndim):
        if periods[ax] is None:
            continue
        pk[ax] = _periodic_knots_dev(kn[ax], periods[ax])
        kuse[ax] = pk[ax][0]
        if functional:  # slopes on the padded tables and knots (_spline.py:924-938 then :1027-1040)
            tabs["f"] = _pad_periodic_dev(tabs["f"], ax, pk[ax][1])
            sk[ax] = pk[ax][0]
            mask |= L.IPX_WRAP_ONLY(ax)
        else:
            perms[ax] = pk[ax][1]
            mask |= L.IPX_PERIODIC(ax)
    for name, src, ax in CHAIN[ndim]:
        tabs[name] = _slopes_dev(sk[ax], tabs[src], method, ax, kwargs)
    return io, ndim, fshape, kn, qd, qshape, params, pk, mask, perms, kuse, sk, tabs


def _jvp_knots_chain(qs, knots, f, knots_dot, f_dot, method, derivative, extrap, period, functional, kwargs):
    """jvp_knots for the methods evaluated from tables (forward mode of _vjp_knots_chain)."""
    io, ndim, fshape, kn, qd, qshape, params, pk, mask, perms, kuse, sk, tabs = _chain_setup(
        qs, knots, f, f, method, derivative, extrap, period, functional, kwargs)
    names = TABLES[ndim]
    lib = L.load()
    suf = _suffix(tabs["f"])
    kd = [None if v is None else _to_dev(v, io.rdt, io.dev).reshape(-1).contiguous() for v in knots_dot]
    # tangents of the lookup knots and of the knots the slopes are taken on
    kd_look, kd_sk = list(kd), list(kd)
    tf = None if f_dot is None else _to_dev(f_dot, io.rdt, io.dev).contiguous()
    for ax in range(ndim):
        if pk[ax] is None:
            continue
        n = fshape[ax]
        src_ix = pk[ax][1].to(torch.int64)[(torch.arange(n + 2, device=io.dev) - 1) % n]
        kd_look[ax] = None if kd[ax] is None else kd[ax][src_ix].contiguous()
        if functional:
            kd_sk[ax] = kd_look[ax]
            if tf is not None:
                tf = _pad_periodic_dev(tf, ax, pk[ax][1])
    tt = {"f": tf if tf is not None else torch.zeros_like(tabs["f"])}
    for name, src, ax in CHAIN[ndim]:
        x = sk[ax]
        n = x.shape[0]
        outer = int(np.prod(tabs[src].shape[:ax], dtype=np.int64))
        inner = int(np.prod(tabs[src].shape[ax + 1:], dtype=np.int64))
        opts, keep = _slope_opts_values(tabs[src], method, ax, kwargs)
        out = torch.empty_like(tabs[src])
        code = _SLOPE_CODE[method]
        L.check(getattr(lib, f"ipx_slopes_jvp_rows_{suf}")(
            _ptr(x), n, _ptr(tabs[src]), _ptr(tabs[name]) if method == "cubic2" else None, _ptr(kd_sk[ax]),
            _ptr(tt[src]), outer, inner, code, C.byref(opts), _ptr(out), _stream()), "ipx_slopes_jvp_rows")
        if method == "cubic2":
            ws = torch.empty(max(int(lib.ipx_slopes_workspace_bytes(code, n, out.element_size())), 16),
                             dtype=torch.uint8, device=out.device)
            L.check(getattr(lib, f"ipx_cubic2_solve_{suf}")(_ptr(x), n, opts.bc_start, opts.bc_end, _ptr(out), outer,
                                                            inner, _ptr(ws), _stream()), "ipx_cubic2_solve")
        tt[name] = out
    pshape = list(tabs["f"].shape)
    narr = (C.c_int32 * ndim)(*[int(pshape[ax]) for ax in range(ndim)])
    tan = torch.empty(qd[0].numel(), dtype=io.tdt, device=io.dev)
    L.check(getattr(lib, f"ipx_jvp_knots_tables_{suf}")(
        ndim, _ptr_array(qd), qd[0].numel(), _ptr_array(kuse), _ptr_array(perms), narr,
        _ptr_array([tabs[n] for n in names]), mask, C.byref(params), 0, _ptr_array(kd_look), _ptr(tan), _stream()),
        "ipx_jvp_knots_tables")
    # the evaluation is linear in the tables: its tangent is the evaluation of the tangent tables, with
    # every extrapolation fill (a constant) replaced by zero
    zp = L.Params()
    for ax in range(ndim):
        a, b = params.axis[ax], zp.axis[ax]
        b.derivative, b.keep_lo, b.keep_hi, b.period = a.derivative, a.keep_lo, a.keep_hi, a.period
        b.fill_lo = b.fill_hi = 0.0
    lin, _ = _eval_dev(ndim, qd, kuse, perms, pshape, [tt[n] for n in names], L.IPX_SOA, 1, L.IPX_CUBIC, mask, zp)
    tan = (tan + lin.reshape(-1)).reshape(qshape)
    return tan if io.want_torch else tan.cpu().numpy()


def vjp_knots(qs, knots, f, cotangent, method="cubic", derivative=0, extrap=False, period=None,
              functional=True, **kwargs):
    """Reverse-mode rule of ``interp{1,2,3}d(q, knots, f, method)`` / ``Interpolator{1,2,3}D(knots, f,
    method)(q)`` with respect to the KNOTS and to ``f``, slope chain included -- what
    ``jax.jacrev(lambda xp: interp1d(x, xp, fp, method))`` computes in the reference
    (tests/test_interpolate.py:542-637).  ``cotangent`` has the shape of the output.  Returns
    ``{"x": ..., ["y", "z",] "f": ...}``.  ``functional``: which of the reference's two periodic
    conventions (slopes on the padded knots, or -- the classes -- on the original ones).
    method = cubic, cardinal, catmull-rom (one kernel: per-query dual numbers over the local knots),
    cubic2 and monotonic (chained: table cotangents, then every link of the slope chain in reverse --
    a transposed tridiagonal solve per cubic2 link)."""
    if method in _CHAIN_KNOT_METHODS:
        return _vjp_knots_chain(qs, knots, f, cotangent, method, derivative, extrap, period, functional, kwargs)
    io, ndim, qd, qshape, kn, kuse, perms, ft, params, mask, code, c, narr = _grad_stencil_setup(
        qs, knots, f, method, derivative, extrap, period, functional, kwargs, cotangent)
    cot = _to_dev(cotangent, io.rdt, io.dev).reshape(-1)
    errorif(cot.numel() != qd[0].numel(), ValueError, "cotangent must have the shape of the output")
    gk = [torch.zeros_like(k) for k in kn]
    gf = torch.zeros_like(ft)
    fn = getattr(L.load(), f"ipx_vjp_stencil_{_suffix(ft)}")
    L.check(fn(ndim, _ptr_array(qd), qd[0].numel(), _ptr_array(kuse), _ptr_array(kn), _ptr_array(perms), narr,
               _ptr(ft), code, c, mask, C.byref(params), 0, _ptr(cot), _ptr_array(gk), _ptr(gf), _stream()),
            "ipx_vjp_stencil")
    out = dict(zip("xyz", gk))
    out["f"] = gf
    return out if io.want_torch else {k: v.cpu().numpy() for k, v in out.items()}


def jvp_knots(qs, knots, f, knots_dot, f_dot=None, method="cubic", derivative=0, extrap=False, period=None,
              functional=True, **kwargs):
    """Forward-mode rule of the same map: the tangent of the output for tangents ``knots_dot``
    (a tuple, entries may be None) and ``f_dot`` (or None)."""
    if method in _CHAIN_KNOT_METHODS:
        return _jvp_knots_chain(qs, knots, f, knots_dot, f_dot, method, derivative, extrap, period, functional,
                                kwargs)
    io, ndim, qd, qshape, kn, kuse, perms, ft, params, mask, code, c, narr = _grad_stencil_setup(
        qs, knots, f, method, derivative, extrap, period, functional, kwargs, f)
    kd = [None if v is None else _to_dev(v, io.rdt, io.dev).reshape(-1) for v in knots_dot]
    fd = None if f_dot is None else _to_dev(f_dot, io.rdt, io.dev).contiguous()
    tan = torch.empty(qd[0].numel(), dtype=io.tdt, device=io.dev)
    fn = getattr(L.load(), f"ipx_jvp_stencil_{_suffix(ft)}")
    L.check(fn(ndim, _ptr_array(qd), qd[0].numel(), _ptr_array(kuse), _ptr_array(kn), _ptr_array(perms), narr,
               _ptr(ft), code, c, mask, C.byref(params), 0, _ptr_array(kd), _ptr(fd), _ptr(tan), _stream()),
            "ipx_jvp_stencil")
    tan = tan.reshape(qshape)
    return tan if io.want_torch else tan.cpu().numpy()


def bin_index(x, q):
    """``clip(searchsorted(x, q, side="right"), 1, len(x)-1)`` on the device -- the interval
    lookup of _spline.py:571 on its own (int32)."""
    want_torch = isinstance(x, torch.Tensor) or isinstance(q, torch.Tensor)
    dev = _device()
    dt = _promote([x, q])
    xt, qt = _to_dev(x, dt, dev), _to_dev(q, dt, dev)
    shape = qt.shape
    qt = qt.reshape(-1)
    idx = torch.empty(qt.numel(), dtype=torch.int32, device=dev)
    lib = L.load()
    fn = getattr(lib, f"ipx_bin_index_{_suffix(xt)}")
    L.check(fn(_ptr(xt), xt.shape[0], _ptr(qt), qt.numel(), _ptr(idx), _stream()), "ipx_bin_index")
    idx = idx.reshape(shape)
    return idx if want_torch else idx.cpu().numpy()


# --------------------------------------------------------------------------- containers
# table name -> slot in the packed record, keyed by the record length (last axis fastest)
_RECORD_SLOT = {
    2: {"f": 0, "fx": 1},
    4: {"f": 0, "fy": 1, "fx": 2, "fxy": 3},
    8: {"f": 0, "fz": 1, "fy": 2, "fyz": 3, "fx": 4, "fxz": 5, "fxy": 6, "fxyz": 7},
}
_FUSED_SLOPES = ("cubic", "cardinal", "catmull-rom", "monotonic", "monotonic-0")


def _slopes_pack_dev(ndim, kn, f, batch, method, kwargs):
    """f -> IPX_AOS records in one launch (ipx_slopes_pack_*), or None where the C ABI says
    IPX_EUNSUPPORTED (1-D, cubic2 / akima, an axis shorter than 3 knots)."""
    if ndim < 2 or method not in _FUSED_SLOPES or any(int(f.shape[ax]) < 3 for ax in range(ndim)):
        return None
    lib = L.load()
    opts = L.SlopeOpts()
    opts.c = float(kwargs.get("c", 0)) if method == "cardinal" else 0.0
    out = torch.empty(tuple(f.shape) + (1 << ndim,), dtype=f.dtype, device=f.device)
    narr = (C.c_int32 * ndim)(*[int(f.shape[ax]) for ax in range(ndim)])
    fn = getattr(lib, f"ipx_slopes_pack_{_suffix(f)}")
    rc = fn(ndim, _ptr_array(list(kn)), narr, _ptr(f), batch, _SLOPE_CODE[method], C.byref(opts), _ptr(out),
            _stream())
    if rc == L.IPX_EUNSUPPORTED:
        return None
    L.check(rc, "ipx_slopes_pack")
    return out


class _RecordTables:
    """The 2^d tables of an interpolator whose records were built in one pass: `f` is the
    caller's array; a slope table is split off the packed records (a strided device copy)
    the first time it is read and kept."""

    def __init__(self, names, f, packed):
        self._names = names
        self._packed = packed
        self._have = {"f": f}

    def __getitem__(self, name):
        if name not in self._have:
            f = self._have["f"]
            slot = _RECORD_SLOT[len(self._names)][name]
            self._have[name] = self._packed.view(*f.shape, len(self._names))[..., slot].contiguous()
        return self._have[name]


class _InterpolatorND:
    """Shared machinery of Interpolator1D/2D/3D (_spline.py:48-441): tables live on the
    device; slopes are precomputed once on the UN-padded tables (:380-393) and, for the
    cubic family, interleaved into the AoS layout the evaluation kernel streams best."""

    _ndim = 0

    def _setup(self, knots, f, method, extrap, period, kwargs):
        ndim = self._ndim
        names = TABLES[ndim]
        kwargs = dict(kwargs)
        self.axis = kwargs.pop("axis", 0)
        given = {n: kwargs.pop(n, None) for n in names[1:]}
        methods = (METHODS_1D, METHODS_2D, METHODS_3D)[ndim - 1]
        fshape = _shape(f)
        _check_shapes(ndim, knots, fshape)
        errorif(method not in methods, ValueError, f"unknown method {method}")
        io = _Inputs((), knots, f)
        kn = [io.coords(k) for k in knots]
        errorif(ndim == 1 and self.axis != 0, NotImplementedError, "only axis=0 is supported")
        self.method = method
        self.extrap = extrap
        self.period = period
        self._io = io
        self._fshape = fshape
        self._knots = kn
        self._f = io.table(f)
        self._given = {n: None if given[n] is None else io.table(given[n]) for n in names[1:]}
        self._kwargs = kwargs
        self._batch = int(np.prod(self._f.shape[ndim:], dtype=np.int64))
        self._pk_cache = {}
        self._packed = None
        self._tabs = None
        self._halo = None
        self._plans = {}  # derivative orders -> call plan of small device-resident calls (_fast_call)
        if (method in _STENCIL_METHODS and self._batch == 1 and not io.complex
                and all(v is None for v in given.values())):
            # linear three-point stencils on a scalar table: evaluate straight from f (one padded
            # copy of f instead of 2^d slope tables); `.derivs` and the packed records are built
            # only if somebody asks for them
            periods = _parse_ndarg(period, ndim)
            pk = [self._periodic(ax, periods[ax]) if periods[ax] is not None else None for ax in range(ndim)]
            per = [p is not None for p in pk]
            # knots that are not sorted inside one period put the one-sided end stencils in the
            # middle of the sorted table: such tables keep the record form
            if all(p is None or p[1] is None for p in pk) and _stencil_fits(
                    [kn[ax].shape[0] + (2 if per[ax] else 0) for ax in range(ndim)], kn[0].element_size()):
                self._halo = _halo_dev(ndim, self._f, per, [None] * ndim, False)
                if ndim == 1:
                    return  # (2-D / 3-D keep the records as well: the faster form depends on the stream)
        self._build_records()

    def _build_records(self):
        """Slope tables on the UN-padded grid (_spline.py:380-393) and, for the cubic family, the
        packed records the tile kernel streams."""
        if self._tabs is not None:
            return
        ndim = self._ndim
        names = TABLES[ndim]
        kn, io, method, kwargs = self._knots, self._io, self.method, self._kwargs
        if all(v is None for v in self._given.values()):
            # nothing supplied: one fused pass f -> packed records; the separate slope tables
            # (`.derivs`) are split off the records only if somebody asks for them
            self._packed = _slopes_pack_dev(ndim, kn, self._f, self._batch, method, kwargs)
        if self._packed is not None:
            self._tabs = _RecordTables(names, self._f, self._packed)
            return
        tabs = {"f": self._f}
        tabs.update(self._given)
        for name, src, ax in CHAIN[ndim]:
            if tabs[name] is None:
                tabs[name] = _slopes_dev(kn[ax], tabs[src], method, ax, kwargs, io.complex)
        self._tabs = tabs
        if _method_class(method) == L.IPX_CUBIC:
            self._packed = _pack_dev(ndim, [tabs[n] for n in names])

    # the reference exposes these as pytree leaves
    @property
    def f(self):
        return self._user(self._f)

    @property
    def derivs(self):
        self._build_records()
        return {n: self._user(self._tabs[n]) for n in TABLES[self._ndim][1:]}

    def _user(self, t):
        if self._io.complex:
            return torch.view_as_complex(t)
        return t

    def _periodic(self, ax, per):
        key = (ax, abs(float(per)))
        if key not in self._pk_cache:
            xpad, perm = _periodic_knots_dev(self._knots[ax], per)
            # knots already sorted inside one period (the usual case): the permutation is the
            # identity and the kernels skip the lookup when handed NULL.  One host sync, at
            # the first call only.
            ident = torch.arange(perm.numel(), dtype=perm.dtype, device=perm.device)
            if bool(torch.equal(perm, ident)):
                perm = None
            self._pk_cache[key] = (xpad, perm)
        return self._pk_cache[key]

    def _fast_call(self, plan, qs, out):
        """A small call with device tensors of the stored type (the README case: 1e4 queries),
        where Python and ctypes -- not the kernel -- are the cost: every argument that does not
        depend on the queries was converted once (`plan`); only the pointers, the count and the
        result are new.  Returns None if the call is not of that kind."""
        q0 = qs[0]
        tdt, dev = plan["tdt"], plan["dev"]
        if type(q0) is not torch.Tensor or q0.dtype != tdt or q0.device != dev or not q0.is_contiguous():
            return None
        shape = q0.shape
        nq = q0.numel()
        if nq == 0 or nq >= _FAST_CALL_MAX:
            return None
        ndim = self._ndim
        qarr = (C.c_void_p * ndim)()
        qarr[0] = q0.data_ptr()
        for k in range(1, ndim):
            q = qs[k]
            if type(q) is not torch.Tensor or q.dtype != tdt or q.device != dev or q.shape != shape or not q.is_contiguous():
                return None
            qarr[k] = q.data_ptr()
        if out is None:
            res = torch.empty(shape, dtype=tdt, device=dev)
        else:
            if (type(out) is not torch.Tensor or out.dtype != tdt or out.device != dev or out.numel() != nq
                    or not out.is_contiguous()):
                return None
            res = out
        a = plan["args"]
        rc = plan["fn"](a[0], qarr, nq, a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], 0, res.data_ptr(), None,
                        stencil_hint, None, _stream())
        if rc != 0:
            return None  # (the general path reports the error, or falls back to the record form)
        return res if out is None else res.view(shape)

    def _make_plan(self, derivative, kuse, ns, mask, params):
        """Plan of `_fast_call` for these derivative orders (evaluation from f, scalar real table)."""
        ndim = self._ndim
        keep = (_ptr_array(kuse), _ptr_array(self._knots), (C.c_int32 * ndim)(*[int(n) for n in ns]), params,
                list(kuse), self._halo, self._packed)
        cten = float(self._kwargs.get("c", 0)) if self.method == "cardinal" else 0.0
        return {
            "tdt": self._io.tdt, "dev": self._f.device, "keep": keep,
            "fn": getattr(L.load(), f"ipx_eval_stencil_{_suffix(self._f)}"),
            "args": (ndim, keep[0], keep[1], keep[2], self._halo.data_ptr(),
                     None if self._packed is None else self._packed.data_ptr(), _STENCIL_METHODS[self.method], cten,
                     mask, C.byref(params)),
        }

    def _call(self, qs, derivative, out=None, return_index=False, presort=False):
        ndim = self._ndim
        st = self._io
        if not return_index and not presort and self._plans:
            plan = self._plans.get(tuple(derivative))
            if plan is not None:
                res = self._fast_call(plan, qs, out)
                if res is not None:
                    return res
        # promotion of the call mirrors interpNd(xq, self.x, self.f, ...)
        io = _Inputs(qs, [np.zeros((), st.xdt)], None, stored=st)
        if io.rdt != st.rdt:
            # queries wider than the stored tables: use the functional form, which promotes
            # everything exactly as the reference does
            self._build_records()
            kw = {n: self._user(self._tabs[n]) for n in TABLES[ndim][1:]}
            if out is not None:
                kw["out"] = out
            if presort:
                kw["presort"] = True
            res = _interp(ndim, qs, [k for k in self._knots], self._user(self._f),
                          self.method, derivative, self.extrap, self.period, kw, return_index)
            if io.want_torch:
                return res
            to_np = lambda t: t.cpu().numpy() if isinstance(t, torch.Tensor) else t
            return tuple(to_np(r) for r in res) if return_index else to_np(res)
        periods = _parse_ndarg(self.period, ndim)
        derivs = _parse_ndarg(derivative, ndim)
        ex = _parse_extrap(self.extrap, ndim)
        mask = 0
        perms = [None] * ndim
        kuse = list(self._knots)
        for ax in range(ndim):
            if periods[ax] is not None:
                kuse[ax], perms[ax] = self._periodic(ax, periods[ax])
                mask |= L.IPX_PERIODIC(ax)
        mclass = _method_class(self.method)
        ns = [self._f.shape[ax] for ax in range(ndim)]
        params = _make_params(ndim, derivs, ex, periods)
        batch = self._batch
        state = {}

        def launch_records(dq, dout, want_idx):
            if "tables" not in state:
                self._build_records()
                state["tables"] = ([self._packed], L.IPX_AOS) if mclass == L.IPX_CUBIC else ([self._f], L.IPX_SOA)
            tables, layout = state["tables"]
            return _eval_dev(ndim, dq, kuse, perms, ns, tables, layout, batch, mclass, mask, params,
                             want_idx, dout)

        def launch(dq, dout, want_idx):
            if self._halo is not None:
                # Interpolator* takes its slopes on the un-padded knots, then wraps: IPX_PERIODIC
                res = _eval_stencil_dev(ndim, dq, kuse, self._knots, ns, self._halo, _STENCIL_METHODS[self.method],
                                        float(self._kwargs.get("c", 0)) if self.method == "cardinal" else 0.0,
                                        mask, params, want_idx, dout, self._packed)
                if res is not None:
                    return res
                self._halo = None  # knot vectors too long for shared memory: record form from now on
                self._plans.clear()
            return launch_records(dq, dout, want_idx)

        if presort:
            plain = _presorted(launch, ndim, kuse, ns, mask, params, batch)
            if self._halo is not None and batch == 1:
                launch = _presorted_stencil(plain, ndim, kuse, self._knots, ns, ns, mask, self._halo,
                                            _STENCIL_METHODS[self.method],
                                            float(self._kwargs.get("c", 0)) if self.method == "cardinal" else 0.0,
                                            mask, params)
            else:
                launch = plain
        res = _execute(io, qs, launch, batch, self._fshape[ndim:], out, return_index)
        if (self._halo is not None and not st.complex and io.want_torch and len(self._plans) < 16
                and all(isinstance(d, (int, np.integer)) for d in derivative)):
            self._plans.setdefault(tuple(derivative), self._make_plan(derivative, kuse, ns, mask, params))
        return res


class Interpolator1D(_InterpolatorND):
    """Convenience class for a 1D interpolated function; same contract as
    interpax.Interpolator1D (_spline.py:48-150).  Two optional extensions of ``__call__``:
    ``out=`` a preallocated result tensor (e.g. pinned host memory for host-resident queries);
    ``presort=True`` visits the queries in table-cell order (device-side radix sort, gather,
    evaluate, scatter) -- same results bit for bit, much faster for random-order queries on
    tables larger than L2."""

    _ndim = 1

    def __init__(self, x, f, method="cubic", extrap=False, period=None, **kwargs):
        self._setup((x,), f, method, extrap, period, kwargs)

    @property
    def x(self):
        return self._knots[0]

    def __call__(self, xq, dx=0, out=None, presort=False):
        return self._call((xq,), (dx,), out, presort=presort)


class Interpolator2D(_InterpolatorND):
    """Same contract as interpax.Interpolator2D (_spline.py:153-278)."""

    _ndim = 2

    def __init__(self, x, y, f, method="cubic", extrap=False, period=None, **kwargs):
        self._setup((x, y), f, method, extrap, period, kwargs)

    @property
    def x(self):
        return self._knots[0]

    @property
    def y(self):
        return self._knots[1]

    def __call__(self, xq, yq, dx=0, dy=0, out=None, presort=False):
        return self._call((xq, yq), (dx, dy), out, presort=presort)


class Interpolator3D(_InterpolatorND):
    """Same contract as interpax.Interpolator3D (_spline.py:281-441)."""

    _ndim = 3

    def __init__(self, x, y, z, f, method="cubic", extrap=False, period=None, **kwargs):
        self._setup((x, y, z), f, method, extrap, period, kwargs)

    @property
    def x(self):
        return self._knots[0]

    @property
    def y(self):
        return self._knots[1]

    @property
    def z(self):
        return self._knots[2]

    def __call__(self, xq, yq, zq, dx=0, dy=0, dz=0, out=None, presort=False):
        return self._call((xq, yq, zq), (dx, dy, dz), out, presort=presort)

"""GPU parity tests of the evaluation straight from f (ipx_eval_stencil_*):
method = cubic / cardinal / catmull-rom on scalar tables, both kernel shapes (one query per lane;
four consecutive queries per lane sharing the table rows), both periodic conventions of the
reference (Interpolator*: slopes on the un-padded knots; interpNd: slopes on the padded knots),
against the CPU oracle on the same seeded inputs.  Bars as in test_gpu_parity.py."""

import numpy as np
import pytest
import torch

import interpax_b200 as ipx
from interpax_b200 import _lib as L
from interpax_b200 import api
from oracle import interpax_np as O
from test_gpu_parity import assert_parity, knots_nonuniform, upcast

pytestmark = pytest.mark.gpu

STENCIL_METHODS = ("cubic", "cardinal", "catmull-rom")


class hint:
    """Force a kernel shape for the calls inside the block."""

    def __init__(self, h):
        self.h = h

    def __enter__(self):
        self.old = api.stencil_hint
        api.stencil_hint = self.h

    def __exit__(self, *exc):
        api.stencil_hint = self.old


def cell_order(knots, qs, periods):
    key = np.zeros(len(qs[0]), dtype=np.int64)
    for k, q, p in zip(knots, qs, periods):
        v = np.mod(q, p) if p is not None else q
        key = key * (len(k) + 3) + np.searchsorted(np.sort(np.mod(k, p)) if p is not None else k, v, side="right")
    return np.argsort(key, kind="stable")


def same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", STENCIL_METHODS)
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_stencil_shapes_vs_oracle(ndim, method, dtype):
    """Every (periodic mask, derivative, extrap) combination below is evaluated by the sparse
    kernel, by the dense kernel on the stream as given (random order: every slot goes through
    the one-query redo path) and by the dense kernel on the cell-sorted stream (rows shared),
    in the class and the functional form.  All three agree bit for bit (a query's result does
    not depend on the kernel shape or on its neighbours) and match the oracle."""
    rng = np.random.default_rng(1000 * ndim + len(method) + (dtype == np.float32))
    ofun = (O.interp1d, O.interp2d, O.interp3d)[ndim - 1]
    gfun = (ipx.interp1d, ipx.interp2d, ipx.interp3d)[ndim - 1]
    ocls = (O.Interpolator1D, O.Interpolator2D, O.Interpolator3D)[ndim - 1]
    gcls = (ipx.Interpolator1D, ipx.Interpolator2D, ipx.Interpolator3D)[ndim - 1]
    kw = {"c": 0.25} if method == "cardinal" else {}
    nq = 3001
    for trial, ns in enumerate([(2, 3, 4), (7, 5, 9), (12, 6, 3)]):
        ns = ns[:ndim]
        uniform = trial == 1
        # trials 0, 1: knots inside [0, period): the classes evaluate periodic axes from f too;
        # trial 2: knots that wrap (the class keeps the record form there, the functional form not)
        los = [float(rng.uniform(-2, 0)) if trial == 2 else 0.0 for _ in range(ndim)]
        his = [lo + float(rng.uniform(1, 4)) for lo in los]
        knots = [np.linspace(lo, hi, n).astype(dtype) if uniform else knots_nonuniform(rng, n, lo, hi, dtype)
                 for lo, hi, n in zip(los, his, ns)]
        if trial == 2 and ns[0] > 6:
            knots[0][4] = knots[0][5]  # a zero-width interval
        f = rng.normal(size=ns).astype(dtype)
        qs = []
        for lo, hi in zip(los, his):
            span = hi - lo
            q = rng.uniform(lo - 0.3 * span, hi + 0.3 * span, nq)
            q[:6] = [lo, hi, lo - 0.1, hi + 0.1, np.nan, lo + 3.7 * span]
            qs.append(q.astype(dtype))
        for mask in range(1 << ndim):
            periods = [(hi - lo) * 1.07 if (mask >> ax) & 1 else None for ax, (lo, hi) in enumerate(zip(los, his))]
            period = None if mask == 0 else (periods[0] if ndim == 1 else tuple(periods))
            deriv = tuple(int(v) for v in rng.integers(0, 4, ndim)) if (mask + trial) % 2 else (0,) * ndim
            extrap = [False, True, (True, 0.5), tuple((False, 2.0) for _ in range(ndim))][(mask + trial) % 4]
            if ndim == 1 and isinstance(extrap, tuple) and isinstance(extrap[0], tuple):
                extrap = extrap[0]
            order = cell_order(knots, qs, periods)
            qsorted = [q[order] for q in qs]
            what = f"nd={ndim} {method} ns={ns} mask={mask} d={deriv} ex={extrap}"
            dfun = deriv[0] if ndim == 1 else deriv
            want_f = ofun(*qs, *knots, f, method, dfun, extrap, period, **kw)
            want_c = ocls(*knots, f, method, extrap, period, **kw)(*qs, *deriv)
            truth_f = truth_c = None
            if dtype == np.float32:
                q64, k64 = [upcast(q) for q in qs], [upcast(k) for k in knots]
                truth_f = ofun(*q64, *k64, upcast(f), method, dfun, extrap, period, **kw)
                truth_c = ocls(*k64, upcast(f), method, extrap, period, **kw)(*q64, *deriv)
            ip = gcls(*knots, f, method, extrap, period, **kw)
            if mask == 0 or trial != 2:
                assert ip._halo is not None, what
            res_f, res_c = {}, {}
            for name, h, stream in (("sparse", L.IPX_STENCIL_SPARSE, qs), ("dense", L.IPX_STENCIL_DENSE, qs),
                                    ("dense-sorted", L.IPX_STENCIL_DENSE, qsorted), ("auto", L.IPX_STENCIL_AUTO, qsorted),
                                    ("column", L.IPX_STENCIL_COLUMN, qs), ("column-sorted", L.IPX_STENCIL_COLUMN, qsorted)):
                with hint(h):
                    gf = gfun(*stream, *knots, f, method, dfun, extrap, period, **kw)
                    gc = ip(*stream, *deriv)
                if stream is qsorted:
                    back = np.empty_like(gf)
                    back[order] = gf
                    gf = back
                    back = np.empty_like(gc)
                    back[order] = gc
                    gc = back
                res_f[name], res_c[name] = gf, gc
                assert_parity(gf, want_f, f"functional {name} {what}", truth_f)
                assert_parity(gc, want_c, f"class {name} {what}", truth_c)
            for name in ("dense", "dense-sorted", "auto", "column", "column-sorted"):
                assert same_bits(res_f[name], res_f["sparse"]), f"functional {name} {what}"
                assert same_bits(res_c[name], res_c["sparse"]), f"class {name} {what}"


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_stencil_indices_and_record_form_agree(ndim):
    """Interval indices from the stencil kernels are bit-exact, and the evaluation from f agrees
    with the record form (slope tables + A_CUBIC contraction) to rounding."""
    rng = np.random.default_rng(50 + ndim)
    ns = [11, 8, 13][:ndim]
    knots = [knots_nonuniform(rng, n, -1.0, 2.0, np.float64) for n in ns]
    f = rng.normal(size=ns)
    nq = 40_001
    qs = [rng.uniform(-1.4, 2.4, nq) for _ in range(ndim)]
    for q, k in zip(qs, knots):
        q[10] = np.nan
        q[11:11 + len(k)] = k
    order = cell_order(knots, qs, [None] * ndim)
    for stream in (qs, [q[order] for q in qs]):
        for h in (L.IPX_STENCIL_SPARSE, L.IPX_STENCIL_DENSE):
            with hint(h):
                res, idx = api._interp(ndim, stream, knots, f, "cubic", 0, True, None, {}, return_index=True)
            for ax in range(ndim):
                assert np.array_equal(idx[ax], O.bin_index(knots[ax], stream[ax]))
            # record form: hand the slope tables in, which selects the gather + contraction kernels
            tabs = {"f": f}
            for name, src, ax in api.CHAIN[ndim]:
                tabs[name] = O.approx_df(knots[ax], tabs[src], "cubic", ax)
            fun = (ipx.interp1d, ipx.interp2d, ipx.interp3d)[ndim - 1]
            rec = fun(*stream, *knots, f, "cubic", 0, True, None, **{n: tabs[n] for n in api.TABLES[ndim][1:]})
            m = ~np.isnan(res)
            assert np.array_equal(np.isnan(rec), ~m)
            np.testing.assert_allclose(res[m], rec[m], rtol=1e-12, atol=1e-12 * np.abs(rec[m]).max())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_stencil_non_finite_table_entries(dtype):
    """NaN / inf in f reach exactly the queries whose 4^d neighbourhood carries a NON-ZERO weight
    on them in the reference's arithmetic -- the dense kernel's zero-padded taps must not spread
    them to neighbours along the last axis (such results are redone by the one-query path)."""
    rng = np.random.default_rng(77)
    ns = (9, 8, 40)
    knots = [np.linspace(0, 1, n).astype(dtype) for n in ns]
    f = rng.normal(size=ns).astype(dtype)
    f[4, 4, 10] = np.nan
    f[2, 6, 25] = np.inf
    f[7, 1, 33] = -np.inf
    nq = 60_000
    qs = [rng.uniform(0, 1, nq).astype(dtype) for _ in range(3)]
    order = cell_order(knots, qs, [None] * 3)
    qsorted = [q[order] for q in qs]
    with hint(L.IPX_STENCIL_SPARSE):
        sparse = ipx.interp3d(*qsorted, *knots, f, "cubic")
    with hint(L.IPX_STENCIL_DENSE):
        dense = ipx.interp3d(*qsorted, *knots, f, "cubic")
    with hint(L.IPX_STENCIL_COLUMN):
        column = ipx.interp3d(*qsorted, *knots, f, "cubic")
    assert np.isnan(sparse).any() and (~np.isnan(sparse)).any()
    assert same_bits(dense, sparse)
    assert same_bits(column, sparse)
    # against the record form of the reference's arithmetic: same set of non-finite results
    sel = np.linspace(0, nq - 1, 8000).astype(np.int64)
    # (float32: compare with the float64 evaluation of the same float32 inputs -- the float32
    # oracle's dense form is itself only good to a few 1e-5 here)
    want = O.interp3d(*[upcast(q[sel]) for q in qsorted], *[upcast(k) for k in knots], upcast(f), "cubic")
    got = sparse[sel].astype(np.float64)
    # every value of the 4^3 neighbourhood enters both forms (a zero weight times NaN or inf is
    # NaN in either), so the same queries come out non-finite
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    fin = np.isfinite(want) & np.isfinite(got)
    rtol = 1e-12 if dtype == np.float64 else 1e-5
    np.testing.assert_allclose(got[fin], want[fin], rtol=rtol, atol=rtol * np.abs(want[fin]).max())


def test_stencil_full_size_c4_dense_vs_sparse():
    """256^3, period on x and y, extrap on z (C4), 2^23 cell-sorted queries: the dense kernel
    (what the benchmark times) against the sparse kernel bit for bit, against the record form
    (Interpolator3D on supplied slope tables) to 1e-12, and against the oracle on a subsample."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(99)
    n = 256
    P = 2 * np.pi
    x = torch.arange(n, dtype=torch.float64, device=dev) * (P / n)
    z = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    f = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
    period = (P, P, None)
    extrap = ((True, True), (True, True), (True, 0.0))
    nq = 1 << 23
    xq = (torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P
    yq = (torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P
    zq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.1 - 0.05
    # cell-sorted, eight queries per cell on average: restrict the stream to 1/16 of the table
    xq = xq.remainder(P / 4)
    yq = yq.remainder(P / 4)
    h = P / n
    key = ((xq / h).floor().long() * n + (yq / h).floor().long()) * (n + 2) + (zq.clamp(0, 1) * (n - 1)).floor().long()
    order = torch.argsort(key)
    xq, yq, zq = xq[order], yq[order], zq[order]
    ip = ipx.Interpolator3D(x, x, z, f, "cubic", extrap, period)
    assert ip._halo is not None
    with hint(L.IPX_STENCIL_DENSE):
        dense = ip(xq, yq, zq)
        dense_dx = ip(xq, yq, zq, 1, 0, 0)
    with hint(L.IPX_STENCIL_SPARSE):
        sparse = ip(xq, yq, zq)
        sparse_dx = ip(xq, yq, zq, 1, 0, 0)
    with hint(L.IPX_STENCIL_COLUMN):
        column = ip(xq, yq, zq)
        column_dx = ip(xq, yq, zq, 1, 0, 0)
        # an unsorted stream through the same shape: every window falls back to the one-query path
        perm = torch.randperm(1 << 18, device=dev, generator=g)
        column_rand = ip(xq[perm], yq[perm], zq[perm])
    assert torch.equal(column, sparse) and torch.equal(column_dx, sparse_dx)
    assert torch.equal(column_rand, sparse[perm])
    auto = ip(xq, yq, zq)
    assert torch.equal(dense, sparse) and torch.equal(dense_dx, sparse_dx) and torch.equal(auto, sparse)
    assert torch.all(dense[zq > 1.0] == 0.0) and torch.isfinite(dense).all()
    ip._build_records()
    rec = ipx.Interpolator3D(x, x, z, f, "cubic", extrap, period, **ip.derivs)
    assert rec._halo is None
    r = rec(xq, yq, zq)
    assert torch.allclose(dense, r, rtol=1e-12, atol=1e-12 * float(r.abs().max()))
    m = 1 << 14
    sel = torch.linspace(0, nq - 1, m, device=dev).long()
    want = O.Interpolator3D(x.cpu().numpy(), x.cpu().numpy(), z.cpu().numpy(), f.cpu().numpy(), "cubic",
                            extrap, period)(xq[sel].cpu().numpy(), yq[sel].cpu().numpy(), zq[sel].cpu().numpy())
    assert_parity(dense[sel].cpu().numpy(), want)


def test_stencil_abi_device_params_every_shape_and_graph_capture():
    """ipx_eval_stencil_f64 through the C ABI with the traced operands (derivative orders, fills,
    periods) in DEVICE memory -- what an XLA custom call hands over -- for every kernel shape,
    against the same call with host operands, bit for bit; and the automatic choice (probe +
    column / sparse candidates, no records) captured in a CUDA graph."""
    import ctypes as C

    lib = L.load()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(4)
    n = 24
    P = 2 * np.pi
    x = torch.arange(n, dtype=torch.float64, device=dev) * (P / n)
    z = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    f = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
    nq = 1 << 17
    xq = (torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P
    yq = (torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P
    zq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.1 - 0.05
    h = P / n
    key = ((xq.remainder(P) / h).floor().long() * n + (yq.remainder(P) / h).floor().long()) * (n + 2) \
        + (zq.clamp(0, 1) * (n - 1)).floor().long()
    order = torch.argsort(key)
    qs = [xq[order].contiguous(), yq[order].contiguous(), zq[order].contiguous()]
    ip = ipx.Interpolator3D(x, x, z, f, "cubic", ((True, True), (True, True), (True, 0.0)), (P, P, None))
    assert ip._halo is not None
    kuse = [ip._periodic(0, P)[0], ip._periodic(1, P)[0], z]
    mask = L.IPX_PERIODIC(0) | L.IPX_PERIODIC(1)
    ns = (C.c_int32 * 3)(n, n, n)
    fn = lib.ipx_eval_stencil_f64
    stream = torch.cuda.current_stream().cuda_stream
    ws = torch.zeros(4, dtype=torch.int32, device=dev)
    for deriv in ((0, 0, 0), (1, 0, 2)):
        p = api._make_params(3, deriv, api._parse_extrap(ip.extrap, 3), (P, P, None))
        pbytes = torch.frombuffer(bytearray(bytes(p)), dtype=torch.uint8).to(dev)
        pdev = C.cast(C.c_void_p(pbytes.data_ptr()), C.POINTER(L.Params))
        want = O.Interpolator3D(x.cpu().numpy(), x.cpu().numpy(), z.cpu().numpy(), f.cpu().numpy(), "cubic",
                                ip.extrap, ip.period)(*[q[:4096].cpu().numpy() for q in qs], *deriv)
        ref = None
        for hint_code in (L.IPX_STENCIL_SPARSE, L.IPX_STENCIL_DENSE, L.IPX_STENCIL_COLUMN, L.IPX_STENCIL_AUTO):
            outs = []
            for params, pod in ((C.byref(p), 0), (pdev, 1)):
                out = torch.full((nq,), -3.0, dtype=torch.float64, device=dev)
                rc = fn(3, api._ptr_array(qs), nq, api._ptr_array(kuse), api._ptr_array([x, x, z]), ns,
                        ip._halo.data_ptr(), None, L.IPX_SLOPE_CUBIC, 0.0, mask, params, pod, out.data_ptr(), None,
                        hint_code, ws.data_ptr() if hint_code == L.IPX_STENCIL_AUTO else None, stream)
                assert rc == L.IPX_OK
                outs.append(out)
            torch.cuda.synchronize()
            assert torch.equal(outs[0], outs[1]), (deriv, hint_code)
            if ref is None:
                ref = outs[0]
                assert_parity(ref[:4096].cpu().numpy(), want)
            assert torch.equal(outs[0], ref), (deriv, hint_code)
    # the automatic choice as a CUDA graph (the probe, both candidates and their skip logic captured)
    p = api._make_params(3, (0, 0, 0), api._parse_extrap(ip.extrap, 3), (P, P, None))
    out_g = torch.zeros(nq, dtype=torch.float64, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    keep = (api._ptr_array(qs), api._ptr_array(kuse), api._ptr_array([x, x, z]))
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            rc = fn(3, keep[0], nq, keep[1], keep[2], ns, ip._halo.data_ptr(), None, L.IPX_SLOPE_CUBIC, 0.0, mask,
                    C.byref(p), 0, out_g.data_ptr(), None, L.IPX_STENCIL_AUTO, ws.data_ptr(),
                    torch.cuda.current_stream().cuda_stream)
            assert rc == L.IPX_OK
    graph.replay()
    torch.cuda.synchronize()
    with hint(L.IPX_STENCIL_SPARSE):
        assert torch.equal(out_g, ip(*qs))


def test_small_calls_reuse_a_call_plan():
    """Interpolator*.__call__ with small device-resident query sets goes through a per-instance
    call plan after the first call (api._fast_call): same results as the general path for every
    derivative order, N-D query shapes, `out=`, and anything the plan does not cover (other dtype,
    NumPy queries, non-contiguous tensors) still takes the general path."""
    dev = torch.device("cuda")
    rng = np.random.default_rng(3)
    xk = np.sort(rng.uniform(0, 1, 30))
    yk = np.linspace(-1, 1, 17)
    f = rng.normal(size=(30, 17))
    ip = ipx.Interpolator2D(torch.from_numpy(xk).to(dev), torch.from_numpy(yk).to(dev), torch.from_numpy(f).to(dev),
                            "cubic", extrap=(False, True))
    xq = torch.from_numpy(rng.uniform(-0.1, 1.1, (7, 50))).to(dev)
    yq = torch.from_numpy(rng.uniform(-1.2, 1.2, (7, 50))).to(dev)
    for d in ((0, 0), (1, 0), (0, 2)):
        first = ip(xq, yq, *d)           # general path; leaves the plan behind
        assert tuple(d) in ip._plans
        second = ip(xq, yq, *d)          # plan
        out = torch.empty(7 * 50, dtype=torch.float64, device=dev)
        third = ip(xq, yq, *d, out=out)  # plan, caller's result tensor
        want = O.Interpolator2D(xk, yk, f, "cubic", (False, True))(xq.cpu().numpy().ravel(), yq.cpu().numpy().ravel(),
                                                                   *d).reshape(7, 50)
        assert second.shape == (7, 50) and third.shape == (7, 50) and third.data_ptr() == out.data_ptr()
        assert same_bits(first.cpu().numpy(), second.cpu().numpy()) and same_bits(first.cpu().numpy(), third.cpu().numpy())
        assert_parity(second.cpu().numpy(), want)
    # not covered by the plan: float32 queries (promotion), NumPy queries, strided tensors
    r32 = ip(xq.float(), yq.float())
    assert r32.dtype == torch.float64
    rnp = ip(xq.cpu().numpy(), yq.cpu().numpy())
    assert isinstance(rnp, torch.Tensor) or isinstance(rnp, np.ndarray)
    rs = ip(xq.t(), yq.t())
    assert rs.shape == (50, 7) and same_bits(rs.t().cpu().numpy(), ip(xq, yq).cpu().numpy())


@pytest.mark.parametrize("ndim", [1, 2, 3])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_presort_folded_into_the_evaluation_is_bitwise_the_plain_call(ndim, dtype):
    """presort=True on the evaluation from f: cell keys + radix sort, then ONE kernel that reads query
    order[k] from the interleaved coordinate records and writes out[order[k]]
    (ipx_eval_stencil_perm_*).  Same bits as the plain call on the one-query-per-lane shape, in the
    caller's order: periodic axes in both conventions, fills, NaN queries, odd counts."""
    from interpax_b200 import _lib as L
    from interpax_b200 import api
    rng = np.random.default_rng(700 + ndim)
    ns = [41, 33, 27][:ndim]
    knots = [np.sort(rng.uniform(0.0, 2.0, n)).astype(dtype) for n in ns]
    f = rng.normal(size=ns).astype(dtype)
    nq = 150_001
    qs = [rng.uniform(-0.4, 2.5, nq).astype(dtype) for _ in range(ndim)]
    qs[0][11] = np.nan
    cls = (ipx.Interpolator1D, ipx.Interpolator2D, ipx.Interpolator3D)[ndim - 1]
    fun = (ipx.interp1d, ipx.interp2d, ipx.interp3d)[ndim - 1]
    old = api.stencil_hint
    try:
        api.stencil_hint = L.IPX_STENCIL_SPARSE
        for method, extrap, period, deriv in (("cubic", (True, 0.5), None, 0), ("cardinal", False, 2.2, 1),
                                              ("catmull-rom", True, 2.2, 0)):
            per = period if ndim == 1 or period is None else (period,) + (None,) * (ndim - 1)
            d = deriv if ndim == 1 else (deriv,) + (0,) * (ndim - 1)
            dd = (deriv,) + (0,) * (ndim - 1)
            kw = {"c": 0.3} if method == "cardinal" else {}
            ip = cls(*knots, f, method, extrap, per, **kw)
            assert ip._halo is not None
            a = ip(*qs, *dd)
            b = ip(*qs, *dd, presort=True)
            assert np.array_equal(np.isnan(a), np.isnan(b))
            m = ~np.isnan(a)
            assert np.array_equal(a[m], b[m]), (method, "class")
            a = fun(*qs, *knots, f, method, d, extrap, per, **kw)
            b = fun(*qs, *knots, f, method, d, extrap, per, presort=True, **kw)
            m = ~np.isnan(a)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[m], b[m]), (method, "functional")
    finally:
        api.stencil_hint = old

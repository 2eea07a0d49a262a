"""GPU parity tests at the sizes BASELINE.json names (C2 and C5; C1, C3, C4 are in
test_gpu_parity.py / test_gpu_stencil.py), and query-shape handling (N-D / broadcast queries).
The oracle sees either the full tables (C2) or sub-blocks of them on which the three-point
stencils give the identical arithmetic (C5: 512^3)."""

import numpy as np
import pytest
import torch

import interpax_b200 as ipx
from oracle import interpax_np as O
from test_gpu_parity import assert_parity, upcast

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ C2: 1e6 knots x 64 columns
@pytest.mark.parametrize("method", ["linear", "cubic2", "monotonic", "akima"])
def test_c2_at_size(method):
    """interp1d method sweep at the C2 shape: 1e6 knots, fp (1e6, 64), one 2^20-query chunk in
    sorted and in random order (the row-streaming kernel), f64, against the oracle on the full
    tables (its slope pass included: cubic2 is a global tridiagonal solve, akima has a global
    threshold)."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1234 + 2)
    n, B = 1_000_000, 64
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    if method in ("monotonic", "akima"):  # the jittered-sorted knots of SURVEY 8(d)
        h = 1.0 / (n - 1)
        x = torch.sort(x + 0.3 * h * (torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1)).values
    fp = torch.randn((n, B), dtype=torch.float64, device=dev, generator=g)
    nq = 1 << 20
    xq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
    xs = torch.sort(xq).values
    ip = ipx.Interpolator1D(x, fp, method, extrap=(True, 0.25))
    r_rand = ip(xq)
    r_sort = ip(xs)
    assert r_rand.shape == (nq, B)
    # the functional form (slopes recomputed inside the call) on a smaller chunk
    m = 1 << 11
    sel = torch.linspace(0, nq - 1, m, device=dev).long()
    r_fun = ipx.interp1d(xq[sel], x, fp, method, 0, (True, 0.25))
    qh = np.concatenate([xq[sel].cpu().numpy(), xs[sel].cpu().numpy()])
    want = O.interp1d(qh, x.cpu().numpy(), fp.cpu().numpy(), method, 0, (True, 0.25))
    assert_parity(r_rand[sel].cpu().numpy(), want[:m], f"C2 {method} random")
    assert_parity(r_sort[sel].cpu().numpy(), want[m:], f"C2 {method} sorted")
    assert_parity(r_fun.cpu().numpy(), want[:m], f"C2 {method} functional")


# ------------------------------------------------------------------ C5: Interpolator3D 512^3, d/dxq
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_c5_at_size(dtype):
    """Interpolator3D on 512^3 (precomputed at construction), derivative (1, 0, 0) -- what
    jit(vmap(grad(interp, argnums=0))) evaluates -- on 2^22 in-range queries, cell-sorted and in
    random order (the device picks the evaluation from f resp. the records), and the value.
    The oracle evaluates sub-blocks of the table: the cubic stencils reach one node to each side,
    so on the cells whose four-node neighbourhood lies inside a block (or ends at a true table
    end) the block reproduces the full-table arithmetic exactly."""
    dev = torch.device("cuda")
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    g = torch.Generator(device=dev).manual_seed(1234 + 5)
    n = 512
    P = 2 * np.pi
    x = (torch.arange(n, device=dev, dtype=torch.float64) * (P / n)).to(tdt)
    z = torch.linspace(0, 1, n, device=dev, dtype=tdt)
    f = torch.randn((n, n, n), dtype=tdt, device=dev, generator=g)
    ip = ipx.Interpolator3D(x, x, z, f, "cubic", True)
    xh, zh = x.cpu().numpy(), z.cpu().numpy()
    nb = 20  # nodes per block and axis
    blocks = [(0, 0, 0), (n - nb, n - nb, n - nb), (137, 301, 0), (250, 0, n - nb), (400, 250, 250)]
    nq = 1 << 22
    per = nq // len(blocks)
    qx, qy, qz, owner = [], [], [], []
    for bi, (bx, by, bz) in enumerate(blocks):
        rng = []
        for b0, k in ((bx, xh), (by, xh), (bz, zh)):
            lo = k[0] if b0 == 0 else k[b0 + 1]
            hi = k[-1] if b0 + nb == n else k[b0 + nb - 2]
            rng.append((float(lo), float(hi)))
        for (lo, hi), dst in zip(rng, (qx, qy, qz)):
            dst.append(torch.rand(per, dtype=torch.float64, device=dev, generator=g) * (hi - lo) * 0.999 + lo)
        owner.append(torch.full((per,), bi, device=dev))
    q = [torch.cat(v).to(tdt) for v in (qx, qy, qz)]
    owner = torch.cat(owner)
    perm = torch.randperm(q[0].numel(), device=dev, generator=g)
    qr = [v[perm] for v in q]  # random order; `q` itself is block-ordered
    h = P / n
    key = ((qr[0] / h).floor().long() * n + (qr[1] / h).floor().long()) * n + (qr[2] * (n - 1)).floor().long()
    order = torch.argsort(key)
    qs = [v[order] for v in qr]  # cell-sorted
    res = {}
    for d in ((1, 0, 0), (0, 0, 0)):
        r_sorted = ip(*qs, *d)
        back = torch.empty_like(r_sorted)
        back[order] = r_sorted
        r_random = ip(*qr, *d)
        # the two orders may be served by different table forms: equal to rounding
        tol = 1e-12 if dtype == np.float64 else 1e-5
        assert torch.allclose(back, r_random, rtol=tol, atol=tol * float(r_random.abs().max()))
        res[d] = (back, r_random)
    own = owner[perm].cpu().numpy()
    m = 1500  # oracle queries per block
    for bi, (bx, by, bz) in enumerate(blocks):
        idx = np.nonzero(own == bi)[0][:m]
        sub = f[bx:bx + nb, by:by + nb, bz:bz + nb].cpu().numpy()
        kx, ky, kz = xh[bx:bx + nb], xh[by:by + nb], zh[bz:bz + nb]
        qh = [v[idx].cpu().numpy() for v in qr]
        o = O.Interpolator3D(kx, ky, kz, sub, "cubic", True)
        o64 = O.Interpolator3D(upcast(kx), upcast(ky), upcast(kz), upcast(sub), "cubic", True) if dtype == np.float32 else None
        for d, (back, r_random) in res.items():
            want = o(*qh, *d)
            truth = o64(*[upcast(v) for v in qh], *d) if o64 is not None else None
            assert_parity(back[idx].cpu().numpy(), want, f"C5 {dtype.__name__} sorted block {bi} d={d}", truth)
            assert_parity(r_random[idx].cpu().numpy(), want, f"C5 {dtype.__name__} random block {bi} d={d}", truth)


# ------------------------------------------------------------------ N-D queries, broadcasting
def _flat(ofun, qs, *rest):
    """The oracle follows the reference, whose cubic branch handles 1-D query arrays only (its
    `(dx * f.T).T` scaling and the `lijk...,li` einsum assume one query axis, _spline.py:1086-1099):
    evaluate the flattened broadcast queries and give the result the broadcast shape."""
    b = np.broadcast_arrays(*[np.asarray(q, dtype=np.float64) for q in qs])
    r = ofun(*[v.reshape(-1) for v in b], *rest)
    return r.reshape(b[0].shape + r.shape[1:])


@pytest.mark.parametrize("method", ["cubic", "linear", "akima"])
def test_nd_queries_broadcast(method):
    """xq, yq, zq of different shapes are broadcast together and the result has shape
    broadcast(xq, yq, zq).shape + f.shape[ndim:] (_spline.py:509-514, 671-677, 904-910): N-D
    query arrays, size-1 axes, Python scalars, with and without trailing batch dims of f."""
    rng = np.random.default_rng(11)
    x, y, z = np.linspace(0, 1, 9), np.linspace(-1, 1, 7), np.linspace(0, 2, 8)
    for batch in ((), (3,), (2, 2)):
        f3 = rng.normal(size=(9, 7, 8) + batch)
        xq, yq, zq = rng.uniform(0, 1, (4, 1, 5)), rng.uniform(-1, 1, (3, 1)), 0.7
        got = ipx.interp3d(xq, yq, zq, x, y, z, f3, method)
        assert got.shape == (4, 3, 5) + batch
        assert_parity(got, _flat(O.interp3d, (xq, yq, zq), x, y, z, f3, method), f"3d {method} {batch}")
        got = ipx.Interpolator3D(x, y, z, f3, method)(xq, yq, zq, 1, 0, 0)
        assert_parity(got, _flat(O.interp3d, (xq, yq, zq), x, y, z, f3, method, (1, 0, 0)),
                      f"3d class {method} {batch}")
        f2 = f3[:, :, 0]
        xq2, yq2 = rng.uniform(0, 1, (2, 3, 1)), rng.uniform(-1, 1, (1, 4))
        got = ipx.interp2d(xq2, yq2, x, y, f2, method)
        assert got.shape == (2, 3, 4) + batch
        assert_parity(got, _flat(O.interp2d, (xq2, yq2), x, y, f2, method), f"2d {method} {batch}")
        f1 = f3[:, 0, 0]
        xq1 = rng.uniform(0, 1, (2, 3, 4))
        got = ipx.interp1d(xq1, x, f1, method)
        assert got.shape == (2, 3, 4) + batch
        assert_parity(got, _flat(O.interp1d, (xq1,), x, f1, method), f"1d {method} {batch}")
    # a 0-d query gives a 0-d result
    s = ipx.interp1d(0.3, x, np.sin(x), "cubic")
    assert np.shape(s) == () and np.isclose(s, O.interp1d(0.3, x, np.sin(x), "cubic"), rtol=1e-12)

"""Driver for the round-2 ncu captures: one process that launches, in order,
  C4 (256^3 f64, period x/y, extrap z, 1e8 cell-sorted queries)  -> eval_tile_kernel on the records
  C4 random order                                               -> stencil_sparse_kernel on the halo table
  C5 (512^3 f64, d/dx, 1.25e8 cell-sorted in-range queries)      -> stencil_sparse_kernel
  C3 (2048^2 f64, 1e8 cell-sorted queries)                       -> stencil_dense_kernel (bulk-copy staging)
each twice (the second launch of each is the one to read)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402

dev = torch.device("cuda")
P = 2 * np.pi
g = torch.Generator(device=dev).manual_seed(1238)
tdt = torch.float64


def sorted3(q, n):
    h = P / n
    key = ((torch.remainder(q[0], P) / h).floor().clamp(0, n - 1).long() * n
           + (torch.remainder(q[1], P) / h).floor().clamp(0, n - 1).long()) * n + (q[2] * (n - 1)).floor().clamp(0, n - 2).long()
    o = torch.argsort(key)
    return [v[o] for v in q]


def run(name, fn):
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1), 3), "ms", flush=True)


which = sys.argv[1:] or ["c4", "c4r", "c5", "c3"]
if "c4" in which or "c4r" in which:
    n, nq = 256, 100_000_000
    x = torch.arange(n, device=dev, dtype=tdt) * (P / n)
    z = torch.linspace(0, 1, n, device=dev, dtype=tdt)
    f = torch.randn((n, n, n), dtype=tdt, device=dev, generator=g)
    ip = ipx.Interpolator3D(x, x, z, f, "cubic", ((True, True), (True, True), (True, 0.0)), (P, P, None))
    q = [(torch.rand(nq, dtype=tdt, device=dev, generator=g) * 3 - 1) * P for _ in range(2)]
    q.append(torch.rand(nq, dtype=tdt, device=dev, generator=g) * 1.1 - 0.05)
    if "c4" in which:
        s = sorted3(q, n)
        run("c4 sorted", lambda: ip(*s))
        del s
    if "c4r" in which:
        run("c4 random", lambda: ip(*q))
    del ip, q, f
    torch.cuda.empty_cache()
if "c5" in which:
    n, nq = 512, 125_000_000
    x = torch.arange(n, device=dev, dtype=tdt) * (P / n)
    z = torch.linspace(0, 1, n, device=dev, dtype=tdt)
    f = torch.randn((n, n, n), dtype=tdt, device=dev, generator=g)
    ip = ipx.Interpolator3D(x, x, z, f, "cubic", False)
    top = P * (n - 1) / n
    q = [torch.rand(nq, dtype=tdt, device=dev, generator=g) * top for _ in range(2)]
    q.append(torch.rand(nq, dtype=tdt, device=dev, generator=g))
    s = sorted3(q, n)
    del q
    run("c5 sorted d/dx", lambda: ip(*s, 1, 0, 0))
    del ip, s, f
    torch.cuda.empty_cache()
if "c3" in which:
    n, nq = 2048, 100_000_000
    x = torch.linspace(0, 1, n, device=dev, dtype=tdt)
    f = torch.randn((n, n), dtype=tdt, device=dev, generator=g)
    ip = ipx.Interpolator2D(x, x, f, "cubic", False)
    q = [torch.rand(nq, dtype=tdt, device=dev, generator=g) * 1.02 - 0.01 for _ in range(2)]
    key = (q[0].clamp(0, 1) * (n - 1)).floor().long() * n + (q[1].clamp(0, 1) * (n - 1)).floor().long()
    o = torch.argsort(key)
    s = [v[o] for v in q]
    run("c3 sorted", lambda: ip(*s))

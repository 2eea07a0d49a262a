"""How much does the query range cost on periodic axes?  The kernels wrap queries within one
period of [0, P) with two adds and sign tests (wrap_one_period) and send anything further out
through the exact remainder (fmod) out of line.  C4 shape, cell-sorted queries, x and y drawn from
[-k P, (k+1) P)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402

dev = torch.device("cuda")
P = 2 * np.pi
n, nq = 256, 100_000_000
g = torch.Generator(device=dev).manual_seed(7)
x = torch.arange(n, device=dev, dtype=torch.float64) * (P / n)
z = torch.linspace(0, 1, n, device=dev, dtype=torch.float64)
f = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
ip = ipx.Interpolator3D(x, x, z, f, "cubic", ((True, True), (True, True), (True, 0.0)), (P, P, None))
base = [torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * P for _ in range(2)]
zq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.1 - 0.05
h = P / n
key = ((base[0] / h).floor().clamp(0, n - 1).long() * n + (base[1] / h).floor().clamp(0, n - 1).long()) * n + (zq * (n - 1)).floor().clamp(0, n - 2).long()
o = torch.argsort(key)
base = [b[o] for b in base]
zq = zq[o]
for k in (0, 1, 2, 50):
    shift = [torch.randint(-k, k + 1, (nq,), device=dev, generator=g).double() * P for _ in range(2)]
    q = [base[0] + shift[0], base[1] + shift[1], zq]
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ip(*q)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"queries within {k} period(s) of [0, P): {np.median(ts[1:]):.3f} ms per 1e8", flush=True)

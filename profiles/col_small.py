import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import interpax_b200 as ipx
from interpax_b200 import _lib as L, api
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(5)
n = 64
P = 2*np.pi
x = torch.arange(n, dtype=torch.float64, device=dev) * (P/n)
z = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
f = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
ip = ipx.Interpolator3D(x, x, z, f, "cubic", ((True, True), (True, True), (True, 0.0)), (P, P, None))
for nq in (1000, 100_000, 2_000_000):
    xq = (torch.rand(nq, dtype=torch.float64, device=dev, generator=g)*3-1)*P
    yq = (torch.rand(nq, dtype=torch.float64, device=dev, generator=g)*3-1)*P
    zq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g)*1.1-0.05
    h = P/n
    key = ((xq.remainder(P)/h).floor().long()*n + (yq.remainder(P)/h).floor().long())*(n+2) + (zq.clamp(0,1)*(n-1)).floor().long()
    o = torch.argsort(key)
    for name, qs in (("sorted", (xq[o], yq[o], zq[o])), ("random", (xq, yq, zq))):
        api.stencil_hint = L.IPX_STENCIL_SPARSE
        a = ip(*qs)
        api.stencil_hint = L.IPX_STENCIL_COLUMN
        b = ip(*qs)
        torch.cuda.synchronize()
        print(nq, name, "equal", bool(torch.equal(a, b)), "maxdiff", float((a-b).abs().max()), flush=True)

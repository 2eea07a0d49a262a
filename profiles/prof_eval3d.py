"""Small driver for ncu: C4-shaped tricubic evaluation (256^3 f64 tables, AoS), a few launches
with cell-sorted then random queries.  Usage: python profiles/prof_eval3d.py [nq] [dtype]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
tdt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else torch.float64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device("cuda")
P = 2 * np.pi
g = torch.Generator(device=dev).manual_seed(1238)
x = (torch.arange(n, device=dev, dtype=torch.float64) * (P / n)).to(tdt)
z = torch.linspace(0, 1, n, device=dev, dtype=tdt)
f = torch.randn((n, n, n), dtype=tdt, device=dev, generator=g)
ip = ipx.Interpolator3D(x, x, z, f, "cubic", ((True, True), (True, True), (True, 0.0)), (P, P, None))
xq = (torch.rand(nq, dtype=tdt, device=dev, generator=g) * 3 - 1) * P
yq = (torch.rand(nq, dtype=tdt, device=dev, generator=g) * 3 - 1) * P
zq = torch.rand(nq, dtype=tdt, device=dev, generator=g) * 1.1 - 0.05
h = P / n
key = ((torch.remainder(xq, P) / h).floor().clamp(0, n - 1).long() * n
       + (torch.remainder(yq, P) / h).floor().clamp(0, n - 1).long()) * n + (zq * (n - 1)).floor().clamp(0, n - 2).long()
o = torch.argsort(key)
sq = [xq[o], yq[o], zq[o]]
torch.cuda.synchronize()
for rep in range(3):
    for qs, name in ((sq, "sorted"), ([xq, yq, zq], "random")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = ip(*qs)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} rep{rep}: {e0.elapsed_time(e1):.3f} ms  {nq / e0.elapsed_time(e1) / 1e6:.2f} Gq/s")

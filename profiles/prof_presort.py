"""Driver for ncu: C4-shaped random-order queries with presort=True (keys, radix sort, gather,
evaluate, scatter).  Usage: python profiles/prof_presort.py [nq]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda")
n = 256
P = 2 * np.pi
g = torch.Generator(device=dev).manual_seed(1238)
x = torch.arange(n, device=dev, dtype=torch.float64) * (P / n)
z = torch.linspace(0, 1, n, device=dev, dtype=torch.float64)
f = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
ip = ipx.Interpolator3D(x, x, z, f, "cubic", ((True, True), (True, True), (True, 0.0)), (P, P, None))
q = [(torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P for _ in range(2)]
q.append(torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.1 - 0.05)
out = torch.empty(nq, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ip(*q, presort=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    print(f"presort rep{rep}: {e0.elapsed_time(e1):.3f} ms  {nq / e0.elapsed_time(e1) / 1e6:.2f} Gq/s")

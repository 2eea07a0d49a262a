"""Small driver for ncu: C3-shaped bicubic evaluation (2048^2 f64, AoS), sorted then random
queries.  Usage: python profiles/prof_eval2d.py [nq]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda")
n = 2048
g = torch.Generator(device=dev).manual_seed(1237)
x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
f = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
ip = ipx.Interpolator2D(x, x, f, "cubic", False)
xq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
yq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
key = (xq.clamp(0, 1) * (n - 1)).floor().long() * n + (yq.clamp(0, 1) * (n - 1)).floor().long()
o = torch.argsort(key)
xs, ys = xq[o], yq[o]
out = torch.empty(nq, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for rep in range(3):
    for qs, name in (((xs, ys), "sorted"), ((xq, yq), "random")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ip(*qs, out=out)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} rep{rep}: {e0.elapsed_time(e1):.3f} ms  {nq / e0.elapsed_time(e1) / 1e6:.2f} Gq/s")

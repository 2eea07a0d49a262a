"""Small driver for ncu: C2-shaped 1-D evaluation (1e6 knots, fp (1e6,64), f64, cubic, AoS),
a few launches with sorted then random queries.  Usage: python profiles/prof_rows.py [nq]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 23
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1236)
n, B = 1_000_000, 64
x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
fp = torch.randn((n, B), dtype=torch.float64, device=dev, generator=g)
ip = ipx.Interpolator1D(x, fp, "cubic")
xq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
xs = torch.sort(xq).values
out = torch.empty((nq, B), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for rep in range(3):
    for q, name in ((xs, "sorted"), (xq, "random")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ip(q, out=out)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} rep{rep}: {e0.elapsed_time(e1):.3f} ms  {nq / e0.elapsed_time(e1) / 1e6:.2f} Gq/s")

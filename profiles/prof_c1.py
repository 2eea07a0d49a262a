"""C1 (README example: 100 knots, 1e4 queries, cubic, f64): microseconds per call of the class and
functional forms with device-resident tensors, and where the host time goes (cProfile)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402

dev = torch.device("cuda")
xp = torch.linspace(0, 2 * np.pi, 100, dtype=torch.float64, device=dev)
xq = torch.linspace(0, 2 * np.pi, 10000, dtype=torch.float64, device=dev)
fp = torch.sin(xp)
ip = ipx.Interpolator1D(xp, fp, "cubic")
out = torch.empty(10000, dtype=torch.float64, device=dev)


def timeit(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("class call            us/call", round(timeit(lambda: ip(xq)), 2))
print("class call, out=      us/call", round(timeit(lambda: ip(xq, out=out)), 2))
print("functional call       us/call", round(timeit(lambda: ipx.interp1d(xq, xp, fp, method="cubic")), 2))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ip(xq, out=out)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        ip(xq, out=out)
print("class call, CUDA graph us/call", round(timeit(lambda: g.replay()), 2))
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    ip(xq, out=out)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

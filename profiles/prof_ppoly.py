"""PPoly evaluation throughput: CubicSpline with n knots, nq queries, trailing batch B.
Usage: python profiles/prof_ppoly.py [n] [nq] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
import interpax_b200 as ipx

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
PEAK = 6540.8
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(3)
x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
y = torch.randn((n,) if B == 1 else (n, B), dtype=torch.float64, device=dev, generator=g)

def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

t_build = timed(lambda: ipx.CubicSpline(x, y, check=False), 3, 1)
cs = ipx.CubicSpline(x, y, check=False)
xq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
xs = torch.sort(xq).values
bytes_ = nq * 8 * (1 + B) + 4 * (n - 1) * B * 8 + n * 8
for name, q in (("sorted", xs), ("random", xq)):
    for nu in (0, 1):
        ms = timed(lambda: cs(q, nu))
        print(f"CubicSpline n={n} B={B} nq={nq:.0e} {name} nu={nu}: {ms:.3f} ms  {nq / ms / 1e6:.2f} Gq/s  "
              f"{bytes_ / ms / 1e6 / PEAK * 100:.1f}% of HBM roofline")
print(f"construction (cubic2 slopes + Hermite -> power basis): {t_build:.3f} ms")

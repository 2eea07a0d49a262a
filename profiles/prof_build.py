"""Time Interpolator3D/2D construction (slopes + pack) : fused one-pass records vs the
per-table chain.  Usage: python profiles/prof_build.py [n3d] [n2d]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from interpax_b200 import api

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

def chain(ndim, kd, fd, method):
    tabs = {"f": fd}
    for name, src, ax in api.CHAIN[ndim]:
        tabs[name] = api._slopes_dev(kd[ax], tabs[src], method, ax, {})
    return api._pack_dev(ndim, [tabs[n] for n in api.TABLES[ndim]])

n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
for dt in (torch.float64, torch.float32):
    for ndim, n in ((3, n3), (2, n2)):
        kd = [torch.linspace(0, 1, n, dtype=dt, device="cuda") for _ in range(ndim)]
        fd = torch.randn(*([n] * ndim), dtype=dt, device="cuda")
        for method in ("cubic", "monotonic"):
            tf = timed(lambda: api._slopes_pack_dev(ndim, kd, fd, 1, method, {}))
            tc = timed(lambda: chain(ndim, kd, fd, method))
            rec = fd.numel() * (1 << ndim) * fd.element_size()
            print(f"{ndim}-D n={n} {str(dt)[6:]} {method:9s}: fused {tf:7.3f} ms ({(rec + fd.numel()*fd.element_size())/tf/1e6:6.0f} GB/s algorithmic)"
                  f"   chain {tc:7.3f} ms   x{tc/tf:.1f}")

"""Timing driver for the evaluation straight from f (ipx_eval_stencil_*): C4 / C5 / C3 shapes,
cell-sorted and random order, the two kernel shapes and the automatic choice, with the record
form (tile kernel) beside it.  Usage: python profiles/prof_stencil.py [c4|c5|c3] [nq] [f64|f32] [what,...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpax_b200 as ipx  # noqa: E402
from interpax_b200 import _lib as L  # noqa: E402
from interpax_b200 import api  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
tdt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
what = sys.argv[4].split(",") if len(sys.argv) > 4 else ["dense", "sparse", "auto", "full", "records"]
reps = int(os.environ.get("REPS", "5"))
dev = torch.device("cuda")
P = 2 * np.pi
g = torch.Generator(device=dev).manual_seed(1238)
if cfg == "c3":
    n = 2048
    x = torch.linspace(0, 1, n, device=dev, dtype=tdt)
    f = torch.randn((n, n), dtype=tdt, device=dev, generator=g)
    mk = lambda **kw: ipx.Interpolator2D(x, x, f, "cubic", False, **kw)
    qs = [torch.rand(nq, dtype=tdt, device=dev, generator=g) * 1.02 - 0.01 for _ in range(2)]
    key = (qs[0].clamp(0, 1) * (n - 1)).floor().long() * n + (qs[1].clamp(0, 1) * (n - 1)).floor().long()
    deriv = [(0, 0), (1, 0)]
else:
    n = 256 if cfg == "c4" else 512
    x = (torch.arange(n, device=dev, dtype=torch.float64) * (P / n)).to(tdt)
    z = torch.linspace(0, 1, n, device=dev, dtype=tdt)
    f = torch.randn((n, n, n), dtype=tdt, device=dev, generator=g)
    if cfg == "c4":
        mk = lambda **kw: ipx.Interpolator3D(x, x, z, f, "cubic", ((True, True), (True, True), (True, 0.0)), (P, P, None), **kw)
        qs = [(torch.rand(nq, dtype=tdt, device=dev, generator=g) * 3 - 1) * P for _ in range(2)]
        qs.append(torch.rand(nq, dtype=tdt, device=dev, generator=g) * 1.1 - 0.05)
        deriv = [(0, 0, 0)]
    else:
        mk = lambda **kw: ipx.Interpolator3D(x, x, z, f, "cubic", False, **kw)
        qs = [torch.rand(nq, dtype=tdt, device=dev, generator=g) * (P * (n - 1) / n) for _ in range(2)]
        qs.append(torch.rand(nq, dtype=tdt, device=dev, generator=g))
        deriv = [(1, 0, 0)]
    h = P / n
    key = ((torch.remainder(qs[0], P) / h).floor().clamp(0, n - 1).long() * n
           + (torch.remainder(qs[1], P) / h).floor().clamp(0, n - 1).long()) * n + (qs[2] * (n - 1)).floor().clamp(0, n - 2).long()
o = torch.argsort(key)
sq = [q[o] for q in qs]
del key, o
ipa = mk()  # halo + records: the device picks per stream ("full")
assert ipa._halo is not None
ipr = mk(**ipa.derivs) if "records" in what else None  # records only
ip = mk()  # halo only
ip._tabs = ip._packed = None
torch.cuda.synchronize()
HINTS = {"dense": L.IPX_STENCIL_DENSE, "sparse": L.IPX_STENCIL_SPARSE, "auto": L.IPX_STENCIL_AUTO,
         "column": L.IPX_STENCIL_COLUMN}
ref = {}
for d in deriv:
    for stream, sname in ((sq, "sorted"), (qs, "random")):
        for w in what:
            obj = ipr if w == "records" else (ipa if w == "full" else ip)
            api.stencil_hint = HINTS.get(w, L.IPX_STENCIL_AUTO)
            if w in ("dense", "column") and sname == "random" and nq > 20_000_000:
                continue  # every slot through the redo path: slow by construction
            ts = []
            for rep in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = obj(*stream, *d)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts[1:]))
            k = (d, sname)
            if k not in ref:
                ref[k] = r
                err = 0.0
            else:
                fin = torch.isfinite(ref[k])
                err = float((r[fin] - ref[k][fin]).abs().max() / ref[k][fin].abs().max())
            print(json.dumps({"cfg": cfg, "nq": nq, "dtype": str(tdt), "deriv": d, "order": sname, "kernel": w,
                              "ms": round(ms, 4), "gqps": round(nq / ms / 1e6, 3), "rel_diff_vs_first": err}), flush=True)
            del r

"""Throughput of the other BASELINE.json configs (C1, C2, C3, C5) on one B200, with a parity
spot-check of each against the oracle.  bench.py measures the headline config (C4); this
script is the evidence for the rest.  Writes one JSON object per config to stdout.

    python profiles/bench_configs.py [c1] [c2] [c3] [c4] [c5]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import interpax_b200 as ipx  # noqa: E402
from oracle import interpax_np as O  # noqa: E402  (checker only)

dev = torch.device("cuda")
PEAK = 6540.8
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timed(fn, steps=10, warmup=3):
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


def parity(got, want, rtol, truth=None):
    """max |got - want| / max |want| against the oracle in the same precision; for float32 also
    the kernel's and the oracle's distance from the float64 evaluation of the same inputs (the
    float32 oracle's own rounding noise can exceed 1e-5 for derivatives on fine grids)."""
    got, want = np.asarray(got), np.asarray(want)
    fin = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    scale = np.max(np.abs(want[fin]))
    err = float(np.max(np.abs(got[fin] - want[fin])) / scale)
    out = {"max_err_over_max_abs": err, "ok": bool(err <= rtol)}
    if truth is not None:
        truth = np.asarray(truth)
        eg = float(np.max(np.abs(got[fin].astype(np.float64) - truth[fin])) / scale)
        eo = float(np.max(np.abs(want[fin].astype(np.float64) - truth[fin])) / scale)
        out.update({"kernel_vs_f64": eg, "oracle_vs_f64": eo, "ok": bool(err <= rtol or eg <= rtol + 2 * eo)})
    return out


def c1():
    xp = np.linspace(0, 2 * np.pi, 100)
    xq = np.linspace(0, 2 * np.pi, 10000)
    fp = np.sin(xp)
    xpt, xqt, fpt = (torch.from_numpy(a).to(dev) for a in (xp, xq, fp))
    ip = ipx.Interpolator1D(xpt, fpt, "cubic")
    ms_f = timed(lambda: ipx.interp1d(xqt, xpt, fpt, method="cubic"), 50, 10)
    ms_c = timed(lambda: ip(xqt), 50, 10)
    got = ipx.interp1d(xq, xp, fp, method="cubic")
    return {"config": "C1 README interp1d cubic, 100 knots, 1e4 queries, f64",
            "us_per_call_functional": ms_f * 1e3, "us_per_call_class": ms_c * 1e3,
            "max_err_vs_sin": float(np.abs(got - np.sin(xq)).max()),
            "parity": parity(got, O.interp1d(xq, xp, fp, method="cubic"), 1e-12)}


def c2():
    out = []
    n, B = 1_000_000, 64
    g = torch.Generator(device=dev).manual_seed(1236)
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    fp = torch.randn((n, B), dtype=torch.float64, device=dev, generator=g)
    nq = 1 << 23  # per chunk; 1e9 queries = 120 such chunks streamed through one 4.3 GB buffer
    xq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
    xs = torch.sort(xq).values
    obuf = torch.empty((nq, B), dtype=torch.float64, device=dev)
    for method in ("linear", "cubic2", "monotonic", "akima"):
        ip = ipx.Interpolator1D(x, fp, method)  # first use also loads the kernels' module
        torch.cuda.synchronize()
        del ip
        t0 = time.perf_counter()
        ip = ipx.Interpolator1D(x, fp, method)
        torch.cuda.synchronize()
        t_build = time.perf_counter() - t0
        ms_r = timed(lambda: ip(xq, out=obuf), 5, 2)
        ms_s = timed(lambda: ip(xs, out=obuf), 5, 2)
        bytes_ = nq * (8 + 8 * B) + (1 if method == "linear" else 2) * n * B * 8 + n * 8
        m = 1 << 12
        got = ip(xq[:m]).cpu().numpy()
        want = O.interp1d(xq[:m].cpu().numpy(), x.cpu().numpy(), fp.cpu().numpy(), method=method)
        out.append({"config": f"C2 interp1d {method}, 1e6 knots, fp (1e6,64), 2^23-query chunk, f64",
                    "queries_per_s_random": nq / ms_r * 1e3, "queries_per_s_sorted": nq / ms_s * 1e3,
                    "roofline_frac_random": bytes_ / (ms_r * 1e-3) / 1e9 / PEAK,
                    "roofline_frac_sorted": bytes_ / (ms_s * 1e-3) / 1e9 / PEAK,
                    "slopes_precompute_s": t_build, "parity": parity(got, want, 1e-12)})
        del ip
    return out


def c3():
    n = 2048
    g = torch.Generator(device=dev).manual_seed(1237)
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    f = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
    ip = ipx.Interpolator2D(x, x, f, "cubic", False)
    nq = 250_000_000
    xq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
    yq = torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.02 - 0.01
    key = (xq.clamp(0, 1) * (n - 1)).floor().long() * n + (yq.clamp(0, 1) * (n - 1)).floor().long()
    o = torch.argsort(key)
    del key
    xs, ys = xq[o], yq[o]
    del o
    obuf = torch.empty(nq, dtype=torch.float64, device=dev)
    res = []
    bytes_ = nq * 24 + 4 * n * n * 8
    for d in ((0, 0), (1, 0)):
        ms_r = timed(lambda: ip(xq, yq, *d, out=obuf), 5, 2)
        ms_s = timed(lambda: ip(xs, ys, *d, out=obuf), 5, 2)
        m = 1 << 12
        got = ip(xq[:m], yq[:m], *d).cpu().numpy()
        want = O.Interpolator2D(x.cpu().numpy(), x.cpu().numpy(), f.cpu().numpy(), "cubic", False)(
            xq[:m].cpu().numpy(), yq[:m].cpu().numpy(), *d)
        res.append({"config": f"C3 bicubic 2048^2, 2.5e8 queries, derivative={d}, extrap=False, f64",
                    "queries_per_s_random": nq / ms_r * 1e3, "queries_per_s_sorted": nq / ms_s * 1e3,
                    "roofline_frac_random": bytes_ / (ms_r * 1e-3) / 1e9 / PEAK,
                    "roofline_frac_sorted": bytes_ / (ms_s * 1e-3) / 1e9 / PEAK,
                    "parity": parity(got, want, 1e-12)})
    return res


def c5():
    res = []
    n = 512
    nq = 125_000_000  # one GPU's share of 1e9 over 8 GPUs
    for tdt, name, rtol in ((torch.float64, "f64", 1e-12), (torch.float32, "f32", 1e-5)):
        g = torch.Generator(device=dev).manual_seed(1239)
        x = torch.linspace(0, 1, n, dtype=tdt, device=dev)
        f = torch.randn((n, n, n), dtype=tdt, device=dev, generator=g)
        ip = ipx.Interpolator3D(x, x, x, f, "cubic")
        q = [torch.rand(nq, dtype=tdt, device=dev, generator=g) for _ in range(3)]
        key = ((q[0] * (n - 1)).floor().long() * n + (q[1] * (n - 1)).floor().long()) * n + (q[2] * (n - 1)).floor().long()
        o = torch.argsort(key)
        del key
        s = [a[o] for a in q]
        del o
        obuf = torch.empty(nq, dtype=tdt, device=dev)
        elem = 8 if name == "f64" else 4
        bytes_ = nq * 4 * elem + 8 * n**3 * elem
        # jit(vmap(grad(interp, argnums=0))) == one derivative-(1,0,0) evaluation
        ms_r = timed(lambda: ip(*q, 1, 0, 0, out=obuf), 3, 2)
        ms_s = timed(lambda: ip(*s, 1, 0, 0, out=obuf), 3, 2)
        m = 1 << 11
        got = ip(*[a[:m] for a in s], 1, 0, 0).cpu().numpy()
        xs = x.cpu().numpy()
        fx = {k: v.cpu().numpy() for k, v in ip.derivs.items()}  # slopes verified separately by the tests
        qs_np = [a[:m].cpu().numpy() for a in s]
        want = O.interp3d(*qs_np, xs, xs, xs, f.cpu().numpy(), "cubic", (1, 0, 0), **fx)
        truth = None
        if name == "f32":
            up = lambda v: v.astype(np.float64)
            truth = O.interp3d(*[up(v) for v in qs_np], up(xs), up(xs), up(xs), up(f.cpu().numpy()), "cubic",
                               (1, 0, 0), **{k: up(v) for k, v in fx.items()})
        res.append({"config": f"C5 Interpolator3D 512^3 d/dxq, 1.25e8 queries (1 GPU share), {name}",
                    "queries_per_s_random": nq / ms_r * 1e3, "queries_per_s_sorted": nq / ms_s * 1e3,
                    "roofline_frac_random": bytes_ / (ms_r * 1e-3) / 1e9 / PEAK,
                    "roofline_frac_sorted": bytes_ / (ms_s * 1e-3) / 1e9 / PEAK,
                    "parity": parity(got, want, rtol, truth)})
        del ip, f, q, s, obuf
        torch.cuda.empty_cache()
    return res


def c4():
    """Around the headline config (bench.py has the class call itself): construction, the
    functional form that rebuilds the slopes on every call, and presorted random order."""
    n, nq = 256, 100_000_000
    P = 2 * np.pi
    g = torch.Generator(device=dev).manual_seed(1238)
    x = torch.arange(n, device=dev, dtype=torch.float64) * (P / n)
    z = torch.linspace(0, 1, n, device=dev, dtype=torch.float64)
    f = torch.randn((n, n, n), dtype=torch.float64, device=dev, generator=g)
    extrap, period = ((True, True), (True, True), (True, 0.0)), (P, P, None)
    ms_build = timed(lambda: ipx.Interpolator3D(x, x, z, f, "cubic", extrap, period), 10, 3)
    ip = ipx.Interpolator3D(x, x, z, f, "cubic", extrap, period)
    q = [(torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 3 - 1) * P for _ in range(2)]
    q.append(torch.rand(nq, dtype=torch.float64, device=dev, generator=g) * 1.1 - 0.05)
    cell = lambda v, per: (torch.remainder(v, per) / per * n).floor().long().clamp(0, n - 1)
    key = (cell(q[0], P) * n + cell(q[1], P)) * n + (q[2].clamp(0, 1) * (n - 1)).floor().long()
    o = torch.argsort(key)
    del key
    s = [a[o] for a in q]
    del o
    obuf = torch.empty(nq, dtype=torch.float64, device=dev)
    ms_cls = timed(lambda: ip(*s, out=obuf), 5, 2)
    ms_fun = timed(lambda: ipx.interp3d(*s, x, x, z, f, "cubic", 0, extrap, period, out=obuf), 5, 2)
    ms_pre = timed(lambda: ip(*q, presort=True, out=obuf), 3, 2)
    ms_rnd = timed(lambda: ip(*q, out=obuf), 3, 2)
    m = 1 << 11
    got = ipx.interp3d(*[a[:m] for a in s], x, x, z, f, "cubic", 0, extrap, period).cpu().numpy()
    want = O.interp3d(*[a[:m].cpu().numpy() for a in s], x.cpu().numpy(), x.cpu().numpy(), z.cpu().numpy(),
                      f.cpu().numpy(), "cubic", 0, extrap, period)
    return {"config": "C4 tricubic 256^3 period+extrap, 1e8 queries, f64: around the headline",
            "construct_ms": ms_build,
            "class_call_sorted_queries_per_s": nq / ms_cls * 1e3,
            "functional_call_sorted_queries_per_s": nq / ms_fun * 1e3,
            "functional_call_ms": ms_fun, "class_call_ms": ms_cls,
            "random_queries_per_s": nq / ms_rnd * 1e3,
            "random_presorted_queries_per_s": nq / ms_pre * 1e3,
            "parity_functional": parity(got, want, 1e-12)}


if __name__ == "__main__":
    which = [a.lower() for a in sys.argv[1:]] or ["c1", "c2", "c3", "c4", "c5"]
    for name in which:
        r = {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}[name]()
        for item in (r if isinstance(r, list) else [r]):
            print(json.dumps(item), flush=True)

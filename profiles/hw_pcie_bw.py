import torch, time, json
torch.cuda.set_device(0)
n = 300_000_000  # 2.4 GB of f64
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.uniform_()
d = torch.empty(n, dtype=torch.float64, device="cuda")
ho = torch.empty(100_000_000, dtype=torch.float64).pin_memory()
do = torch.randn(100_000_000, dtype=torch.float64, device="cuda")
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t = timed(lambda: d.copy_(h, non_blocking=True))
print(json.dumps({"what": "H2D 2.4GB one copy", "ms": t * 1e3, "GBs": 2.4 / t}))
for chunk_mb in (8, 32, 96, 256):
    c = chunk_mb * (1 << 20) // 8
    def f():
        for s in range(0, n, c):
            d[s:s + c].copy_(h[s:s + c], non_blocking=True)
    t = timed(f)
    print(json.dumps({"what": f"H2D chunks {chunk_mb}MB", "ms": t * 1e3, "GBs": 2.4 / t}))
s2 = torch.cuda.Stream()
def duplex():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        ho.copy_(do, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
t = timed(duplex)
print(json.dumps({"what": "H2D 2.4GB + D2H 0.8GB concurrently", "ms": t * 1e3, "GBs_h2d": 2.4 / t}))
t = timed(lambda: ho.copy_(do, non_blocking=True))
print(json.dumps({"what": "D2H 0.8GB", "ms": t * 1e3, "GBs": 0.8 / t}))

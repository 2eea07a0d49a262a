# Random 8-byte gathers from a working set of W MB: where does the L2 stop holding it?
import torch, json
torch.cuda.set_device(0)
n = 1 << 27
g = torch.Generator(device="cuda"); g.manual_seed(1)
out = torch.empty(n, dtype=torch.float64, device="cuda")
for W in (16, 32, 40, 48, 56, 64, 72, 80, 96, 112, 128, 160, 256):
    m = W * (1 << 20) // 8
    buf = torch.randn(m, dtype=torch.float64, device="cuda")
    idx = torch.randint(0, m, (n,), device="cuda", generator=g)
    for _ in range(2):
        torch.index_select(buf, 0, idx, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        torch.index_select(buf, 0, idx, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"W_MB": W, "ms": round(ms, 3), "Gsector_per_s": round(n / ms / 1e6, 2)}), flush=True)

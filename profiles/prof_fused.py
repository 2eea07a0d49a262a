"""One fused slopes+pack launch (3-D f64 cubic, n^3) for ncu.  Usage: prof_fused.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpax_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kd = [torch.linspace(0, 1, n, dtype=torch.float64, device="cuda") for _ in range(3)]
fd = torch.randn(n, n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    out = api._slopes_pack_dev(3, kd, fd, 1, "cubic", {})
torch.cuda.synchronize()
print(float(out.view(-1)[:8].sum()))

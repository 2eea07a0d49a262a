"""Per-SASS-instruction stall samples of one kernel from `ncu --page source --csv`: prints the
instructions with the most samples for each stall reason.  Usage:
  python profiles/stall_profile.py <ncu-rep> [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, b = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        b = {"rows": []}
        blocks.append(b)
    elif b is not None:
        b["rows"].append(r)
h, d = blocks[0]["rows"][0], blocks[0]["rows"][1:]
iA, iS, iI, iSm = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(float(r[iSm] or 0) for r in d)
print("total samples", tot)
for i in stall:
    s = sum(float(r[i] or 0) for r in d)
    if s / tot < 0.04:
        continue
    print(f"== {h[i]}: {s / tot * 100:.1f}% of samples")
    for r in sorted(d, key=lambda r: -float(r[i] or 0))[:topn]:
        print(f"   {float(r[i] or 0) / tot * 100:5.2f}%  {r[iA][-5:]}  {r[iS][:90]}")

"""Attribute executed warp instructions of one kernel to CUDA source lines.
Joins `ncu --page source --csv` (per-SASS-instruction counts) with `nvdisasm --print-line-info`
of the cubin inside the .so (needs -lineinfo).  Usage:
  python profiles/line_profile.py <ncu-rep> <mangled-kernel-substring> [min_pct]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "interpax_b200", "libipx.so")
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
sass = ""
for f in os.listdir(td):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, f)], capture_output=True, text=True).stdout
        if kname in out:
            sass = out
            break
lines = sass.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
addr2 = {}
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if "inlined at" in m.group(3):
            continue  # keep the innermost location (first annotation)
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        addr2[int(m.group(1), 16)] = (cur, m.group(2))
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
iA, iI, iSm = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
base = min(int(r[iA], 16) for r in d if r[iA])
per = collections.Counter()
samp = collections.Counter()
tot = 0.0
for r in d:
    if not r[iA]:
        continue
    off = int(r[iA], 16) - base
    n = float(r[iI] or 0)
    loc = addr2.get(off, (("?", 0), ""))[0]
    per[loc] += n
    samp[loc] += float(r[iSm] or 0)
    tot += n
stot = sum(samp.values()) or 1
srcs = {}
print(f"total warp instructions {tot:.0f}")
for loc, n in sorted(per.items(), key=lambda kv: -kv[1]):
    if n / tot * 100 < min_pct:
        continue
    fn, ln = loc
    if fn not in srcs:
        p = os.path.join(root, "interpax_b200", "csrc", fn)
        srcs[fn] = open(p).read().split("\n") if os.path.exists(p) else []
    text = srcs[fn][ln - 1].strip() if 0 < ln <= len(srcs[fn]) else ""
    print(f"{n / tot * 100:5.1f}% instr {samp[loc] / stot * 100:5.1f}% samples  {fn}:{ln:<4d} {text[:100]}")

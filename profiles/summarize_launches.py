"""Kernel-time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python profiles/summarize_launches.py gpurun_out/launches_r01.csv > profiles/launches_r01.txt
Launches shorter than 20 us are counted apart: with the device-side choice of kernel (ipx_stencil.cu)
every candidate kernel of a call is enqueued and the ones not chosen return at once."""
import collections
import csv
import re
import statistics
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
iK, iV, iU = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
t, n, runs, skips = collections.Counter(), collections.Counter(), collections.defaultdict(list), collections.Counter()
for r in rows[1:]:
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)
    k = re.sub(r"\(.*", "", r[iK])[:100]
    t[k] += v
    n[k] += 1
    if v < 0.02:
        skips[k] += 1
    else:
        runs[k].append(v)
tot = sum(t.values())
print(f"# total kernel time {tot:.3f} ms over {sum(n.values())} launches (cold-cache, serialised: compare shares)")
print("#   total ms  share  launches (of which < 20 us)  median ms of the others  kernel")
for k, v in t.most_common(18):
    med = statistics.median(runs[k]) if runs[k] else 0.0
    print(f"{v:10.3f} ms {v / tot * 100:5.1f}% n={n[k]:4d} ({skips[k]:3d} short) median {med:8.3f} ms  {k}")

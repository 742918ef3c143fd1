"""Condense an `ncu -i X.ncu-rep --page source --csv` dump of ONE kernel into code regions of similar execution count.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python profiles/source_regions.py src.csv [kernel_index]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)
k = kernels[which]
hdr = k["hdr"]
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
iw, ie, ismp = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("# Samples")
data = [(int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[iw] or 0), int(r[ie] or 0), int(r[ismp] or 0)) for r in k["rows"]]
base = data[0][0]
tot = sum(d[2] for d in data)
print(k["name"][:80])
print("instructions", tot, "shared wavefronts", sum(d[3] for d in data), "excessive", sum(d[4] for d in data), "samples", sum(d[5] for d in data))
segs, prev = [], None
for a, s, ex, w, e, smp in data:
    if prev is None or abs(ex - prev) > 0.25 * max(prev, 1):
        segs.append([a - base, a - base, 0, 0, 0, 0, s])
    g = segs[-1]
    g[1] = a - base
    g[2] += 1
    g[3] += ex
    g[4] += w
    g[5] += smp
    prev = ex
for g in segs:
    if g[3] > 0.004 * tot:
        print(f"{g[0]:6x}-{g[1]:6x} n={g[2]:4d} inst={g[3]/1e6:7.2f}M ({100*g[3]/tot:4.1f}%) per-inst={g[3]/g[2]/1e3:8.1f}k wavefronts={g[4]/1e6:6.2f}M samples={g[5]:5d}  {g[6][:48]}")

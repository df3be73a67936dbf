"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel. usage: launch_summary.py <csv> [title]"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
t, n = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[iu].replace("second", "s").replace("nsecond", "ns").replace("usecond", "us").replace("msecond", "ms")]
    k = r[ik].split("(")[0][:90]
    t[k] += v; n[k] += 1
tot = sum(t.values())
print((sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print(f"total {tot:.1f} ms over {sum(n.values())} launches")
for k, v in t.most_common(16):
    print(f"  {v:10.3f} ms  {100 * v / tot:5.1f}%  x{n[k]:4d}  {k}")

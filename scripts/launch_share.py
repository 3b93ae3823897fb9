"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel: share of the summed kernel time,
launch count, average duration.  usage: launch_share.py launches.csv [top N]"""
import collections, csv, re, sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h, rows = rows[hdr], rows[hdr + 1:]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)      # -> microseconds
    name = re.sub(r"\(.*", "", r[ki]).replace("void (anonymous namespace)::", "").replace("void <unnamed>::", "")
    name = name.replace("<unnamed>::", "").replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (serialised, cold caches, unlocked clocks)")
print("| share | launches | avg us | kernel |\n|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"| {100 * t / tot:5.1f} % | {c} | {t / c:.1f} | `{n[:90]}` |")

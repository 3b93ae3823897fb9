"""Top stall lines of an .ncu-rep (source page): python scripts/ncu_top.py file.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second"]
for r in rows[2:]:
    print("kernel:", r[h.index("Kernel Name")][:90])
    for i, n in enumerate(h):
        if n in want: print(f"  {n:70s} {rows[1][i]:10s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {n: i for i, n in enumerate(hdr)}; data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']]) for r in data)
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
agg = {n: sum(int(r[idx[n]]) for r in data) for n in stalls}
print("total samples", tot, "by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:N]:
    s = int(r[idx['# Samples']])
    st = sorted(((int(r[idx[n]]), n) for n in stalls), reverse=True)[:2]
    print(f"{s:6d} {100*s/tot:5.1f}% {r[idx['Source']].strip()[:64]:64s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")

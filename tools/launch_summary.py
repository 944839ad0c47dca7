"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import csv, re, sys, collections
path, out = sys.argv[1], sys.argv[2]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3))
agg = collections.OrderedDict()
for name, us in rows:
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"\(.*$", "", short)
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += us
total = sum(a[1] for a in agg.values())
with open(out, "w") as f:
    f.write(f"# launch list summary: {path.split('/')[-1]} ({len(rows)} launches, {total/1e3:.2f} ms summed, serialised + cold-cache per ncu)\n\n")
    f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {us/1e3:.3f} | {us/n:.1f} | {100*us/total:.1f}% |\n")
print(open(out).read())

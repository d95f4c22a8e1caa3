"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:64]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, n, t, 100 * t / tot, t / n))
print("\ntotal %.1f us over %d launches" % (tot, sum(n for n, _ in agg.values())))

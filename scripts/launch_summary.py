"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {v/1e3:10.3f} {100*v/T:6.1f}% {v/cnt[k]:10.1f}")
print(f"{'TOTAL':60s} {sum(cnt.values()):8d} {T/1e3:10.3f}")

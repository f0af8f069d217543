"""DRAM bytes per walker of one local-energy pass from an ncu launch list that carries
gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum per launch:
python scripts/hbm_from_launches.py launches.csv <walkers covered by the listed launches> out.json [chunk_walkers] [git_sha]"""
import collections, csv, json, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
walkers = float(sys.argv[2])
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
    if name.startswith("void cutlass") or name.startswith("void at::"):
        continue                      # the cuBLAS DGEMM peak probe and torch glue of bench.py are not part of the pass
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v      # -> us
        cnt[name] += 1
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    per[name][m] += v
tot_b = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in per.values())
tot_t = sum(d["gpu__time_duration.sum"] for d in per.values())
out = {"walkers": walkers, "chunk_walkers": int(sys.argv[4]) if len(sys.argv) > 4 else None,
       "git_sha": sys.argv[5] if len(sys.argv) > 5 else None, "source_launch_list": sys.argv[1], "dram_bytes_per_walker": tot_b / walkers, "kernel_us_per_walker": tot_t / walkers,
       "kernels": {k: {"launches": cnt[k], "us": d["gpu__time_duration.sum"],
                       "dram_bytes": d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"],
                       "gb_per_s": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / max(d["gpu__time_duration.sum"], 1e-9) / 1e3}
                   for k, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])}}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(f"DRAM bytes per walker {out['dram_bytes_per_walker'] / 1e6:.1f} MB, kernel time per walker {out['kernel_us_per_walker']:.1f} us")

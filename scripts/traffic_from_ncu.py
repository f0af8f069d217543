"""Per-launch DRAM bytes of every distinct kernel of an `ncu --set full` report, with provenance:
python scripts/traffic_from_ncu.py rep.ncu-rep out.json <chunk_walkers> <git_sha>"""
import csv, io, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
kern = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(CUtensorMap")[0].split("(const")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")
    rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_write.sum"]]]
    t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    tu = units[col["gpu__time_duration.sum"]]
    t_us = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tu, 1.0)
    k = kern.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "us": 0.0,
                               "tensor_pipe_active_pct": [], "grid": r[col["launch__grid_size"]]})
    k["launches"] += 1; k["dram_bytes"] += rd + wr; k["us"] += t_us
    key = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    if key in col and r[col[key]]:
        k["tensor_pipe_active_pct"].append(float(r[col[key]].replace(",", "")))
res = {"chunk_walkers": int(sys.argv[3]) if len(sys.argv) > 3 else None, "git_sha": sys.argv[4] if len(sys.argv) > 4 else None,
       "source": sys.argv[1], "kernels": {}}
for n, k in kern.items():
    tp = k.pop("tensor_pipe_active_pct")
    res["kernels"][n] = {"launches": k["launches"], "dram_bytes_per_launch": k["dram_bytes"] / k["launches"],
                         "us_per_launch": k["us"] / k["launches"], "grid": k["grid"],
                         "tensor_pipe_active_pct": sum(tp) / len(tp) if tp else None}
json.dump(res, open(sys.argv[2], "w"), indent=1)
for n, v in sorted(res["kernels"].items(), key=lambda kv: -kv[1]["us_per_launch"] * kv[1]["launches"]):
    print(f"{n[:60]:60s} {v['launches']:3d} x {v['us_per_launch']:9.1f} us  {v['dram_bytes_per_launch']/1e6:9.1f} MB/launch  tensor {v['tensor_pipe_active_pct']}")

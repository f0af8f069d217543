"""One block of headline metrics per distinct kernel of an ncu report: python scripts/ncu_summary.py rep.ncu-rep > summary.txt"""
import csv, io, subprocess, sys
KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
seen = set()
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key = name.split("(")[0]
    if key in seen:
        continue
    seen.add(key)
    print("---")
    print(f"  Kernel Name [] = {name[:130]}")
    for k in KEYS:
        if k in col:
            print(f"  {k} [{units[col[k]]}] = {r[col[k]]}")

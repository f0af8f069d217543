"""Per-instruction hot spots of one kernel of an ncu report: python scripts/ncu_hot.py rep.ncu-rep <kernel-substring> [min-share]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
for b in blocks:
    name = b[0]
    if pat not in name:
        continue
    rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
    hdr = rows[0]; cols = {h: i for i, h in enumerate(hdr)}
    iS, isamp, ie = cols["Source"], cols["# Samples"], cols["Instructions Executed"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[1:]:
        if len(r) <= ie: continue
        try: data.append((int(r[isamp] or 0), int(r[ie] or 0), r[iS].strip(), r))
        except ValueError: pass
    tot = sum(d[0] for d in data)
    print("==", name[:120], "samples", tot, "instr", len(data), "executed", sum(d[1] for d in data))
    for i, d in enumerate(data):
        if d[0] > thr * tot:
            top = sorted([(int(d[3][cols[s]] or 0), s) for s in stalls], reverse=True)[:2]
            print(f"{i:5d} {100*d[0]/tot:5.1f}% x{d[1]:9d} {d[2][:64]:64s} {top}")
    blk = 100
    print("   per-100-instruction sample share:", " ".join(f"{b0}:{100*sum(d[0] for d in data[b0:b0+blk])/max(tot,1):.0f}" for b0 in range(0, len(data), blk)))

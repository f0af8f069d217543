"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05.mma = UTCIMMA / UTCHMMA...,
tcgen05.ld = LDTM, TMA = UTMALDG, tcgen05.commit = UTCBAR, fp64 tensor = DMMA, DSMEM st.async = STAS / SYNCS):
python scripts/sass_summary.py deepsolid_b200/libdeepsolid_b200.so > profiles/sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "deepsolid_b200/libdeepsolid_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = ["UTCIMMA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "DMMA", "STAS", "SYNCS", "UCGABAR", "REDG", "ATOMG", "RED.", "ATOM"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        cur = re.sub(r"\(.*", "", name)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    ins = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not ins:
        continue
    counts[cur]["_instr"] += 1
    op = ins.group(1)
    for p in PAT:
        if op.startswith(p):
            counts[cur][p] += 1; total[p] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a); git " +
      subprocess.run(["git", "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip())
print("# totals: " + "  ".join(f"{k}={v}" for k, v in total.items()))
hdr = ["UTCIMMA", "LDTM", "UTMALDG", "UTCBAR", "DMMA", "STAS", "UCGABAR", "SYNCS"]
print(f"{'kernel':72s} {'instr':>6s} " + " ".join(f"{h:>8s}" for h in hdr))
for k, c in counts.items():
    if any(c[h] for h in hdr):
        print(f"{k[:72]:72s} {c['_instr']:6d} " + " ".join(f"{c[h]:8d}" for h in hdr))

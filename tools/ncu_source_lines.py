"""Dev tool: executed warp instructions per source line from one .ncu-rep (needs -lineinfo + --import-source on).
usage: python tests/ncu_source_lines.py report.ncu-rep out.csv"""
import csv, subprocess, sys, collections, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
if not raw.strip():
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
for k, r in enumerate(rows):
    if any("Instructions Executed" in c for c in r):
        hdr, body = r, rows[k + 1:]
        break
if hdr is None:
    open(out, "w").write("no source page\n" + raw[:2000])
    sys.exit(0)
def col(name):
    for i, h in enumerate(hdr):
        if h.strip() == name:
            return i
    return None
ci = col("# Instructions Executed") if col("# Instructions Executed") is not None else col("Instructions Executed")
cs = col("Source")
cl = None
for i, h in enumerate(hdr):
    if "Location" in h or h.strip() in ("File", "Line"):
        cl = i
cthr = col("Thread Instructions Executed")
agg = collections.OrderedDict()
total = 0
for r in body:
    if ci is None or len(r) <= ci:
        continue
    try:
        n = int(float(r[ci].replace(",", "") or 0))
    except ValueError:
        continue
    key = (r[cl] if cl is not None and len(r) > cl else "") + " | " + (r[cs][:90] if cs is not None and len(r) > cs else "")
    t = 0
    if cthr is not None and len(r) > cthr:
        try:
            t = int(float(r[cthr].replace(",", "") or 0))
        except ValueError:
            t = 0
    a = agg.setdefault(key, [0, 0])
    a[0] += n
    a[1] += t
    total += n
with open(out, "w") as f:
    f.write("header," + "|".join(hdr) + "\n")
    f.write(f"total_warp_instr,{total}\n")
    for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:150]:
        f.write(f"{n},{100.0*n/max(total,1):.2f}%,{(t/n if n else 0):.1f},{key}\n")

"""Dev tool: executed warp instructions / samples of lav2_kernel split by phase, from one .ncu-rep.
Phases are told apart by the average number of active threads per SASS instruction and by marker opcodes:
usage: python tools/ncu_phase_split.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci, ct, cs, csamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > ct]
tot = sum(int(r[ci]) for r in body)
samp = sum(int(r[csamp]) for r in body)
print("total warp instr", tot, "samples", samp)
# contiguous regions: print cumulative table every ~40 instructions with avg lanes to eyeball phases
acc_i = acc_t = acc_s = 0
start = 0
for k, r in enumerate(body):
    n, t, s = int(r[ci]), int(r[ct]), int(r[csamp])
    acc_i += n; acc_t += t; acc_s += s
    if (k + 1) % 40 == 0 or k == len(body) - 1:
        print(f"sass {start:5d}-{k:5d}: {acc_i/1e6:9.1f} M warp-instr ({100*acc_i/tot:5.1f} %)  samples {100*acc_s/max(samp,1):5.1f} %  lanes {acc_t/max(acc_i,1):5.1f}   {body[start][cs].strip()[:40]}")
        acc_i = acc_t = acc_s = 0
        start = k + 1

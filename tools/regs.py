"""Dev tool: registers / spills per kernel from fractalshark_b200/csrc/build.log.  usage: python tools/regs.py [substring ...]"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(root, "fractalshark_b200", "csrc", "build.log")).read().split("\n")
out, name, extra = [], None, {}
for ln in log:
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        name = m.group(1)
    if "spill" in ln and name:
        extra[name] = ln.strip()
    m = re.search(r"Used (\d+) registers.*", ln)
    if m and name:
        out.append((name, ln.strip()))
dem = subprocess.run(["c++filt"] + [n for n, _ in out], capture_output=True, text=True).stdout.split("\n")
for (n, l), d in zip(out, dem):
    if all(k in d for k in sys.argv[1:]):
        print(d[:150], "|", l[14:], "|", extra.get(n, ""))

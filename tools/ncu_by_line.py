#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel launch in an ncu report.

    [NCU_KERNEL=<demangled substring>] python tools/ncu_by_line.py <report.ncu-rep> <mangled-kernel-substring> [launch-index] [top]

Joins `ncu --page source --csv` (per SASS instruction: executed count, stall samples) with
`nvdisasm -g` of the library's cubin (source line per SASS instruction, needs -lineinfo) by instruction order."""
import csv, os, re, subprocess, sys, tempfile
from collections import Counter

rep, kern = sys.argv[1], sys.argv[2]
kern_ncu = os.environ.get("NCU_KERNEL", kern)       # demangled name in the report when it differs (templates)
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "vapor_b200", "csrc", "libvapor_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("api.")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for l in dis:
    if l.startswith("\t.section\t.text."):
        on = kern in l
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"\s*/\*[0-9a-f]{4,6}\*/", l):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
secs, name = [], ""
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1]
    elif len(r) > 10 and r[0] == "Address":
        secs.append({"name": name, "hdr": r, "data": []})
    elif len(r) > 10 and secs:
        secs[-1]["data"].append(r)
secs = [s for s in secs if kern_ncu in s["name"]]
s = secs[which]
hdr, data = s["hdr"], s["data"]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
assert len(data) == len(lines), (len(data), len(lines))
ci, cs = Counter(), Counter()
for r, ln in zip(data, lines):
    ci[ln] += int(r[iI]); cs[ln] += int(r[iS])
ti, ts = sum(ci.values()), sum(cs.values())
print(f"launch {which} of {len(secs)}: {ti} warp instructions, {ts} samples")
src = {}
for ln, c in ci.most_common(top):
    f = ln[0] if ln else "?"
    if ln and f not in src:
        p = os.path.join(root, "vapor_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src.get(f, [])[ln[1] - 1].strip()[:90] if ln and src.get(f) and ln[1] <= len(src[f]) else ""
    print(f"{100 * c / ti:5.1f}% inst {100 * cs[ln] / max(ts, 1):5.1f}% smp  {f}:{ln[1] if ln else 0:<4d} {text}")

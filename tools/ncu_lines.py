#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel: joins the SASS page of an .ncu-rep (ncu --page source --csv)
with the line table of the matching cubin (nvdisasm --print-line-info), by instruction order.
usage: ncu_lines.py report.ncu-rep object.o kernel_regex [kernel_mangled_substring]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, obj, kre = sys.argv[1:4]
sub = sys.argv[4] if len(sys.argv) > 4 else kre
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]
ix, sx = hdr.index("Instructions Executed"), hdr.index("Source")
stx = hdr.index("Warp Stall Sampling (All Samples)")
sass = []
for r in rows[h + 1:]:
    if len(r) > ix and r[0].startswith("0x"):
        sass.append((r[sx].strip(), int(r[ix] or 0), int(r[stx] or 0)))
    elif len(r) > 1 and r[0] == "Kernel Name" and sass:
        break          # only the first launch
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = sub in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = int(m.group(2))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} profiled", file=sys.stderr)
agg, stall = {}, {}
for (src, n, st), l in zip(sass, lines):
    agg[l] = agg.get(l, 0) + n
    stall[l] = stall.get(l, 0) + st
tot = sum(agg.values()) or 1
stot = sum(stall.values()) or 1
src_lines = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", os.path.basename(obj).replace(".o", ".cu"))).read().splitlines()
print(f"total warp instructions {tot}, stall samples {stot}")
for l, n in sorted(agg.items(), key=lambda kv: -kv[1])[:int(os.environ.get("TOP", 25))]:
    txt = src_lines[l - 1].strip()[:100] if l and l <= len(src_lines) else "?"
    print(f"{n:11d} {100 * n / tot:5.1f}%  stall {100 * stall[l] / stot:5.1f}%  L{l}: {txt}")

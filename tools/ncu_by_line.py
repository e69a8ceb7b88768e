#!/usr/bin/env python3
"""Executed warp instructions and stall samples per CUDA source line of one kernel, by joining the
SASS page of an ncu report with nvdisasm's line info of the same build.

    python tools/ncu_by_line.py REPORT.ncu-rep KERNEL_MANGLED_SUBSTRING [SO]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, ksub = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "sitrack_b200", "libsitrack_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = None
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    m = re.search(r"^\.text\.(\S*%s\S*):\n(.*?)(?=^\s*//-+ \.|\Z)" % re.escape(ksub), txt, flags=re.M | re.S)
    if m and "k_advect" in m.group(1) and ".text." not in m.group(2)[:50]:
        body = m.group(2)
        cur = ("?", 0)
        lines = []
        for ln in body.split("\n"):
            mm = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                lines.append(cur)
        print("kernel:", m.group(1), "instructions:", len(lines))
        break
assert lines, "kernel not found in " + so
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if h is not None:
            break
        h = {k: i for i, k in enumerate(r)}
        continue
    if h is None or len(r) < len(h):
        continue
    try:
        data.append((int(r[h["Instructions Executed"]]), int(r[h["# Samples"]] or 0), r[h["Source"]]))
    except ValueError:
        pass
assert len(data) == len(lines), (len(data), len(lines))
byline = collections.defaultdict(lambda: [0, 0])
for (ie, smp, _), key in zip(data, lines):
    byline[key][0] += ie
    byline[key][1] += smp
tot_i = sum(v[0] for v in byline.values())
tot_s = sum(v[1] for v in byline.values())
nwarps = max(d[0] for d in data)
print("total warp instr %d (%.0f per warp), samples %d" % (tot_i, tot_i / nwarps, tot_s))
print("%-28s %10s %7s %7s" % ("file:line", "instr/warp", "%instr", "%stall"))
for key, v in sorted(byline.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if v[0] / tot_i > 0.004 or v[1] / max(tot_s, 1) > 0.006:
        print("%-28s %10.1f %7.1f %7.1f" % ("%s:%d" % key, v[0] / nwarps, 100 * v[0] / tot_i, 100 * v[1] / max(tot_s, 1)))

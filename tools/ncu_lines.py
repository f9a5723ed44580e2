"""Joins an `ncu --page source --csv` SASS dump with nvdisasm line info: per source line (with inline call chain
collapsed to the innermost bt_* source line) -> instructions executed, stall samples.  Usage:
    python tools/ncu_lines.py <report.ncu-rep> <cubin> [top_n]
"""
import csv
import collections
import re
import subprocess
import sys

rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ins = rows[2:]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines = []
cur = None
infn = False
for l in dis.splitlines():
    mm = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if mm:
        cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]+\*/", l):
        lines.append(cur)
assert len(lines) >= len(ins), (len(lines), len(ins))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for k, r in enumerate(ins):
    key = lines[k]
    n = int(r[ci["Instructions Executed"]] or 0)
    s = int(r[ci["# Samples"]] or 0)
    t = int(r[ci["Thread Instructions Executed"]] or 0)
    agg[key][0] += n; agg[key][1] += s; agg[key][2] += t
    tot[0] += n; tot[1] += s
print("total inst", tot[0], "samples", tot[1])
byfile = collections.defaultdict(lambda: [0, 0])
for (k, v) in agg.items():
    byfile[k[0] if k else None][0] += v[0]; byfile[k[0] if k else None][1] += v[1]
print({k: (round(100 * v[0] / tot[0], 1), round(100 * v[1] / tot[1], 1)) for k, v in byfile.items()})
srcs = {}
for (k, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if k and k[0] not in srcs:
        try:
            srcs[k[0]] = open("/root/repo/brax_tracking_b200/csrc/" + k[0]).read().splitlines()
        except Exception:
            srcs[k[0]] = []
    text = srcs[k[0]][k[1] - 1].strip()[:90] if k and len(srcs.get(k[0], [])) >= k[1] else ""
    print(f"{100*v[0]/tot[0]:5.1f}% inst {100*v[1]/tot[1]:5.1f}% samp  thr/inst {v[2]/max(v[0],1):4.1f}  {k}  {text}")

"""Per-phase (BtEnv member function) share of instructions and stall samples from an ncu report.
    python tools/ncu_phases.py <report.ncu-rep> <cubin>"""
import collections
import csv
import re
import subprocess
import sys

rep, cubin = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}; ins = rows[2:]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# function ranges in bt_impl.h
impl = open(sys.argv[3] if len(sys.argv) > 3 else "/root/repo/brax_tracking_b200/csrc/bt_impl.h").read().splitlines()
funcs = []
for n, l in enumerate(impl, 1):
    mm = re.match(r"\s+(?:static )?(?:template <[^>]*>\s*)?BT_DEV\s+[\w:<>\*&\s]+?\s+(\w+)\(", l)
    if mm and l.startswith("  ") and not l.startswith("    "):
        funcs.append((n, mm.group(1)))
def func_of(line):
    name = "?"
    for n, f in funcs:
        if n <= line: name = f
        else: break
    return name
locs = []
cur = None
for l in dis.splitlines():
    if "//## File" in l:
        chain = re.findall(r'"([^"]*)", line (\d+)', l)
        cur = [(f.split("/")[-1], int(n)) for f, n in chain]
        continue
    if re.match(r"\s*/\*[0-9a-f]+\*/", l):
        locs.append(cur)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
last_key = "other"
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.defaultdict(lambda: collections.Counter())
for k, r in enumerate(ins):
    chain = locs[k] or []
    # phases: outermost bt_impl.h frames, from the outside in
    impl_frames = [func_of(n) for f, n in chain if f == "bt_impl.h"]
    prog = [n for f, n in chain if f == "bt_programs.h"]
    key = impl_frames[-1] if impl_frames else ("programs" if prog else None)
    if key is None:   # helper inlined from bt_math.h / intrinsics: attribute to the enclosing function in address order
        key = last_key
    last_key = key
    # a second-level key: e.g. solve called from solve_constraints
    n = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0); t = int(r[ci["Thread Instructions Executed"]] or 0)
    inner = impl_frames[0] if impl_frames else key
    for kk in {("outer", key), ("inner", inner)}:
        agg[kk][0] += n; agg[kk][1] += s; agg[kk][2] += t
    tot[0] += n; tot[1] += s
    for c in stall_cols:
        v = int(r[ci[c]] or 0)
        if v: stalls[inner][c] += v
print("total warp-inst", tot[0], "samples", tot[1])
for kind in ("outer", "inner"):
    print(f"--- by {kind} function")
    for (kk, v) in sorted(((k, v) for k, v in agg.items() if k[0] == kind), key=lambda kv: -kv[1][1]):
        top = ", ".join(f"{c[6:]} {100*n/max(v[1],1):.0f}%" for c, n in stalls[kk[1]].most_common(3)) if kind == "inner" else ""
        print(f"{kk[1]:20s} inst {100*v[0]/tot[0]:5.1f}%  time(samples) {100*v[1]/tot[1]:5.1f}%  thr/inst {v[2]/max(v[0],1):4.1f}  {top}")

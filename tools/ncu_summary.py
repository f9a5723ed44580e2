"""Turns gpurun_out/prof_<tag>.ncu-rep (ncu --set full capture of one bt_k_step launch, tools/gpu_cycle.sh) into the tracked
summaries under profiles/: per-phase shares, hot source lines, selected raw metrics, the ncu details page, and
profiles/traffic.json (DRAM bytes of the launch, consumed by bench.py as roofline.traffic).

    python tools/ncu_summary.py <tag> [model=rodent] [variant=3_1]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
model = sys.argv[2] if len(sys.argv) > 2 else "rodent"
variant = sys.argv[3] if len(sys.argv) > 3 else "3_1"
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_{model}.ncu-rep")
if not os.path.exists(rep):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
obj = os.path.join(ROOT, "brax_tracking_b200", "build", f"tu_step_{variant}.o")
tmp = "/tmp/bt_cubin"
os.makedirs(tmp, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cubin = os.path.join(tmp, f"tu_step_{variant}.sm_100a.cubin")
out = lambda name: os.path.join(ROOT, "profiles", f"{tag}_{model}_step_kernel_{name}.txt")


def run(cmd, path):
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(path, "w") as f:
        f.write(r.stdout)


run([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), rep, cubin], out("phases"))
run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, cubin, "60"], out("lines"))
run(["ncu", "-i", rep, "--page", "details"], out("details"))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
]
with open(out("metrics"), "w") as f:
    for k in KEYS:
        if k in m:
            f.write(f"{k} [{m[k][0]}] = {m[k][1]}\n")


def to_bytes(k):
    u, v = m[k]
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


tr = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
tp = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(tp)) if os.path.exists(tp) else {}
t[model] = tr
t.setdefault("sources", {})[model] = (f"profiles/{tag}_{model}_*: ncu --set full --clock-control none, bt_k_step_{variant}, "
                                       "dram__bytes_read.sum + dram__bytes_write.sum of one launch")
json.dump(t, open(tp, "w"))
print("traffic", tr, "->", tp)

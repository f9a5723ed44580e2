"""Measured parity numbers of the CUDA path against the oracle, per model and per horizon:
    python tools/parity_report.py > profiles/<tag>_parity_report.json        (on a GPU box)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import common  # noqa: E402
import parity_cases as pc  # noqa: E402
from backends import CudaBackend  # noqa: E402

out = {}
for name in ("rodent", "fly_free", "fly_tethered", "rodent_pair"):
    ep = None if name == "rodent" else 12
    b = CudaBackend(common.setup(name)[3])
    bt = CudaBackend(common.setup(name, ep)[3]) if ep else b
    n = 8 if name == "rodent_pair" else 16
    tf = pc.check_teacher_forced(bt, name, N=n, T=100 if name == "rodent" else 40, episode_length=ep)
    h = pc.check_physics_1_10_100(b, name, N=n)
    out[name] = {
        "teacher_forced_one_control_step": {k: (round(v, 9) if isinstance(v, float) else v) for k, v in tf.items()},
        "free_running_median_abs_error_vs_float64_oracle": {
            f"{label}_{t}_steps": {"qpos_cuda": e, "qpos_oracle_f32": e32, "qvel_cuda": v, "qvel_oracle_f32": v32}
            for (label, t), (e, e32, v, v32) in h.items()},
    }
print(json.dumps(out, indent=1))

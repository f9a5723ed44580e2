"""Tiny driver for compute-sanitizer (racecheck / memcheck): one forward + one wrapped step on 4 rodent envs."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import common, parity_cases as pc
from backends import CudaBackend
name = sys.argv[1] if len(sys.argv) > 1 else "rodent"
m, cfg, clip, tables = common.setup(name)
b = CudaBackend(tables)
st, ctrl = pc.random_states(m, 4)
full, cdist, niter = b.forward_debug(st, ctrl, 0)
print("forward ok", niter)
s2, out = b.reset(common.jax_keys(4))
first = {k: v.copy() for k, v in s2.items()}
b.step(s2, out, first, out["obs"].copy(), out["info_i"].copy(), common.actions(1, 4, m.nu)[0])
torch.cuda.synchronize()
print("step ok", out["reward"])

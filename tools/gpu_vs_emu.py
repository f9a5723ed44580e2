"""Debug helper: per-region comparison of the CUDA scratch dump against the host emulation (same fp32 program).
   python tools/gpu_vs_emu.py make   (CPU: writes gpurun_out/emu_dump.npz)      python tools/gpu_vs_emu.py check   (GPU)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import common, parity_cases as pc
name = sys.argv[2] if len(sys.argv) > 2 else "rodent"
m, cfg, clip, tables = common.setup(name)
N = 16
st, ctrl = pc.random_states(m, N)
stops = (1, 2, 4, 5, 6, 0)
path = os.path.join(ROOT, "tools", "emu_dump.npz")
if sys.argv[1] == "make":
    from backends import EmuBackend
    b = EmuBackend(tables)
    out = {}
    for s in stops:
        sc, cd, ni = b.forward_debug(st, ctrl, s)
        out[f"sc{s}"] = sc; out[f"cd{s}"] = cd; out[f"ni{s}"] = ni
    np.savez_compressed(path, **out)
else:
    from backends import CudaBackend
    b = CudaBackend(tables)
    ref = np.load(path)
    regions = [k[2:] for k in tables if k.startswith("o_")]
    offs = sorted((int(tables["o_" + r][0]), r) for r in regions)
    for s in stops:
        sc, cd, ni = b.forward_debug(st, ctrl, s)
        print(f"--- stop {s}  niter gpu {ni.tolist()} emu {ref[f'ni{s}'].tolist()}")
        for i, (o, r) in enumerate(offs):
            end = offs[i + 1][0] if i + 1 < len(offs) else sc.shape[1]
            a, e = sc[:, o:end], ref[f"sc{s}"][:, o:end]
            if a.size == 0: continue
            d = np.abs(a - e).max(1); sca = np.abs(e).max() + 1e-30
            bad = np.where(d > 1e-3 * sca)[0]
            print(f"   {r:12s} [{o:5d},{end:5d}) maxabs {d.max():.3e} scale {sca:.3e} bad envs {bad.tolist()}")

#!/usr/bin/env python
"""CPU baseline of the hot path, the REAL thing first:

  1. `mjx_available()`: prepend `baseline/_ref/` (git-ignored slot for a reference install with its dependencies) to
     `sys.path` and try `import jax, mujoco, mujoco.mjx` with `JAX_PLATFORMS=cpu`.  If that works, `run_mjx()` times
     `mjx.step x n_frames` (the reference's `pipeline_step`, /root/reference/envs/fruitfly.py:500, with the solver options of
     envs/rodent.py:66-73) for the rodent on the host cores, on the same seeded inputs as the GPU run.
  2. Otherwise (today: none of jax / mujoco / brax is installable in this image) `run_port()` times the repo's C
     restatement (`oracle/`), labelled "port".

`tests/test_mjx_pin.py` uses the same loader to pin the oracle against real MJX the moment the packages exist, and
`bench.py --impl reference` prefers arm 1.  The MJCF assets belong to the reference checkout: they are looked up under
$BT_REFERENCE_ROOT, baseline/_ref/reference, /root/reference.

    python baseline/run_cpu_baseline.py [--envs 16] [--steps 100]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

_REF = os.path.join(HERE, "_ref")


def reference_root():
    for p in (os.environ.get("BT_REFERENCE_ROOT"), os.path.join(_REF, "reference"), "/root/reference"):
        if p and os.path.exists(os.path.join(p, "assets", "rodent.xml")):
            return p
    return None


def mjx_available():
    """-> (jax, mujoco, mjx) or None.  Never raises."""
    if os.path.isdir(_REF) and _REF not in sys.path:
        sys.path.insert(0, _REF)
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    try:
        import jax
        import mujoco
        from mujoco import mjx
        return jax, mujoco, mjx
    except Exception:
        return None


def mjx_model(mujoco, mjx, xml_path, iterations=4, ls_iterations=4):
    """The options the reference env constructors apply (envs/rodent.py:66-73, envs/fruitfly.py:71-78)."""
    mjm = mujoco.MjModel.from_xml_path(xml_path)
    mjm.opt.solver = mujoco.mjtSolver.mjSOL_CG
    mjm.opt.iterations = iterations
    mjm.opt.ls_iterations = ls_iterations
    mjm.opt.jacobian = 0  # dense
    return mjm, mjx.put_model(mjm)


def mjx_pipeline_fn(jax, mjx, mx, n_frames, with_act):
    """jit(vmap(pipeline_step)): (qpos, qvel, act, qacc_warmstart, ctrl) -> the same + xpos after n_frames x mjx.step."""
    def one(qpos, qvel, act, warm, ctrl):
        d = mjx.make_data(mx)
        d = d.replace(qpos=qpos, qvel=qvel, qacc_warmstart=warm, ctrl=ctrl)
        if with_act:
            d = d.replace(act=act)
        if n_frames == 0:
            d = mjx.forward(mx, d)
        for _ in range(n_frames):
            d = mjx.step(mx, d)
        return d.qpos, d.qvel, d.act, (d.qacc if n_frames == 0 else d.qacc_warmstart), d.xpos
    return jax.jit(jax.vmap(one))


def run_mjx(n_envs=16, n_steps=100, n_frames=5, seed=1):
    """BASELINE.json configs[0]: rodent.xml (unscaled when dm_control's rescale is absent), n_envs x n_steps control steps."""
    mods = mjx_available()
    root = reference_root()
    if mods is None or root is None:
        return None
    jax, mujoco, mjx = mods
    mjm, mx = mjx_model(mujoco, mjx, os.path.join(root, "assets", "rodent.xml"))
    fn = mjx_pipeline_fn(jax, mjx, mx, n_frames, mjm.na > 0)
    rng = np.random.default_rng(seed)
    qpos = np.tile(mjm.qpos0, (n_envs, 1)).astype(np.float32)
    qpos[:, 7:] += rng.uniform(-0.05, 0.05, (n_envs, mjm.nq - 7)).astype(np.float32)
    qvel = np.zeros((n_envs, mjm.nv), np.float32)
    act = np.zeros((n_envs, mjm.na), np.float32)
    warm = np.zeros((n_envs, mjm.nv), np.float32)
    acts = np.tanh(rng.standard_normal((n_steps + 1, n_envs, mjm.nu))).astype(np.float32)
    out = fn(qpos, qvel, act, warm, acts[0])
    jax.block_until_ready(out)                       # compile + warm-up
    t0 = time.perf_counter()
    for t in range(n_steps):
        out = fn(out[0], out[1], out[2], out[3], acts[t + 1])
    jax.block_until_ready(out)
    dt = time.perf_counter() - t0
    return dict(kind="reference", value=n_envs * n_steps / dt, wall_s=dt, cores=os.cpu_count() or 1,
                sample=f"{n_envs} envs x {n_steps} control steps, mujoco.mjx {getattr(mujoco, '__version__', '?')} on the JAX CPU backend "
                       f"(jax {jax.__version__}), rodent.xml, CG {mjm.opt.iterations}x{mjm.opt.ls_iterations}, physics only")


def run_port(model="rodent", n_envs=16, n_steps=100, threads=None):
    """The oracle port: C restatement of mjx.step x n_frames + numpy env layer, on `threads` host threads."""
    import common
    import env_oracle
    import oracle as oracle_mod
    threads = threads or (os.cpu_count() or 1)
    m, cfg, clip, _ = common.setup(model)
    o = oracle_mod.Oracle(m, np.float32)
    eo = env_oracle.EnvOracle(o, clip, cfg, dtype=np.float32)
    s = eo.reset(common.jax_keys(n_envs))
    acts = common.actions(n_steps + 1, n_envs, m.nu, seed=1)
    s = eo.step(s, acts[0])
    t0 = time.perf_counter()
    for t in range(n_steps):
        s = eo.step(s, acts[t + 1])
    dt = time.perf_counter() - t0
    return dict(kind="port", value=n_envs * n_steps / dt, wall_s=dt, cores=min(threads, n_envs),
                sample=f"{n_envs} envs x {n_steps} control steps, oracle port (C restatement of mjx.step, float32) + numpy env layer")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16)
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    r = run_mjx(a.envs, a.steps)
    if r is None:
        r = run_port("rodent", a.envs, a.steps)
        r["note"] = "jax / mujoco.mjx not importable (baseline/_ref absent): oracle restatement -- not MJX"
    r["unit"] = "env-steps/s"
    print(json.dumps(r))


if __name__ == "__main__":
    main()

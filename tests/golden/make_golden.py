"""Generates tests/golden/<model>_golden.npz with the float64 oracle (the reference itself is not runnable here:
SURVEY.md F3).  Teacher forcing is not needed: the vectors are short (8 control steps) and only used to freeze the
oracle and to give the product a fixed, committed target (tests/test_golden.py).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import common  # noqa: E402

EPISODE = 6
for name, N, T in (("rodent", 8, 8), ("fly_free", 8, 8), ("fly_tethered", 8, 8), ("rodent_pair", 4, 6)):
    m, cfg, clip, _ = common.setup(name, EPISODE)
    _, eo = common.oracles(name, np.float64, EPISODE)
    keys = common.jax_keys(N, seed=21)
    acts = common.actions(T, N, m.nu, seed=22, scale=0.3)
    s = eo.reset(keys)
    out = dict(keys=keys, actions=acts, episode_length=np.int32(EPISODE), reset_cur_frame=s["info"]["cur_frame"], reset_obs=s["obs"],
               reset_qpos=s["pipeline_state"]["qpos"], reset_qvel=s["pipeline_state"]["qvel"])
    done, cur, rew, steps, trunc, obs0 = [], [], [], [], [], []
    for t in range(T):
        s = eo.step(s, acts[t])
        done.append(s["done"]); cur.append(s["info"]["cur_frame"]); rew.append(s["reward"]); steps.append(s["info"]["steps"])
        trunc.append(s["info"]["truncation"])
        if t == 0:
            out["step1_obs"] = s["obs"]; out["step1_qpos"] = s["pipeline_state"]["qpos"]; out["step1_qvel"] = s["pipeline_state"]["qvel"]
    out.update(done=np.array(done), cur_frame=np.array(cur), reward=np.array(rew), steps=np.array(steps), truncation=np.array(trunc),
               final_qpos=s["pipeline_state"]["qpos"])
    np.savez_compressed(os.path.join(HERE, f"{name}_golden.npz"), **out)
    print(name, "done flags:", int(np.sum(done)), "reward[0]:", rew[0][:3])

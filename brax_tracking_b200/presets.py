"""Named environment presets = compiled asset + the reference's dataset config + the seeded synthetic clip
(SURVEY.md section 8d).  Used by bench.py, train.py and the tests; none of this is in the timed path."""
from __future__ import annotations

import functools

from . import assets, clips, configs

ENV_ARGS = dict(rodent=configs.RODENT_ENV_ARGS, fly_free=configs.FLY_FREEJNT_ENV_ARGS, fly_tethered=configs.FLY_ENV_ARGS,
                # BASELINE.json configs[3]: rodent_pair.xml, both animals tracked, 12 opened inter-animal capsule pairs
                rodent_pair=configs.RODENT_PAIR_ENV_ARGS)


@functools.lru_cache(maxsize=None)
def load(name: str):
    """-> (mjcf.Model, env_args, clip dict)"""
    m = assets.load_model(name)
    args = ENV_ARGS[name]
    return m, args, clips.synthetic_clip(m, args["free_jnt"]).as_dict()


def make_env(name: str, device: int = 0, **overrides):
    from . import envs
    m, args, clip = load(name)
    return envs.TrackingEnv(m, clip, dict(args, **overrides), device=device)

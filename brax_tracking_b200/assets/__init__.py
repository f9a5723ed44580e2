"""Pre-compiled model constants (``<name>.npz``) for the reference's MJCF assets.

The MJCF files themselves belong to the reference checkout and are not redistributed here; the flat
``mjModel``-style tables produced from them by ``mjcf.compile_mjcf`` are (tools/compile_assets.py is the
generating script).  ``load_model`` restores a ``mjcf.Model``; users with their own MJCF call
``mjcf.compile_mjcf`` directly.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .. import mjcf

_DIR = os.path.dirname(os.path.abspath(__file__))
_SCALARS = ("nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "ntendon", "nM", "timestep", "density", "viscosity", "cone",
            "impratio", "tolerance", "ls_tolerance", "iterations", "ls_iterations", "meaninertia")

MODELS = ("rodent", "fly_free", "fly_tethered", "rodent_pair")


def save_model(m: mjcf.Model, path: str) -> None:
    meta = {k: (int(getattr(m, k)) if isinstance(getattr(m, k), (int, np.integer)) else float(getattr(m, k))) for k in _SCALARS}
    meta["gravity"] = [float(x) for x in m.gravity]
    meta["names"] = m.names
    np.savez_compressed(path, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **m.a)


def load_model(name_or_path: str) -> mjcf.Model:
    path = name_or_path if os.path.exists(name_or_path) else os.path.join(_DIR, name_or_path + ".npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no compiled model '{name_or_path}' (known: {MODELS})")
    z = np.load(path)
    meta = json.loads(bytes(z["__meta__"]).decode())
    m = mjcf.Model()
    for k in _SCALARS:
        setattr(m, k, meta[k])
    m.gravity = np.array(meta["gravity"])
    m.names = meta["names"]
    m.a = {k: z[k] for k in z.files if k != "__meta__"}
    return m

"""Reference-clip container and the synthetic clip used by tests and benchmarks.

``ReferenceClip`` mirrors /root/reference/preprocessing/preprocess.py:23-41 (time-major ``[T, ...]`` arrays).
No mocap data ships with the reference, so benchmarks and parity tests run on a seeded
synthetic clip with the same shapes, built like ``process_clip`` builds a real one
(preprocess.py:99-141): per-frame forward kinematics for ``body_positions`` and finite differences
(quaternion log for the root) for the velocities (preprocess.py:207-230).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np

from . import mjcf


@dataclass
class ReferenceClip:
    position: np.ndarray            # [T, 3]   (models with A free roots: [T, 3 A], the roots side by side)
    quaternion: np.ndarray          # [T, 4]   ([T, 4 A])
    joints: np.ndarray              # [T, nq-7]  (tethered: [T, nq]; A roots: [T, nq - 7 A])
    body_positions: np.ndarray      # [T, nbody, 3]
    velocity: np.ndarray            # [T, 3]
    angular_velocity: np.ndarray    # [T, 3]
    joints_velocity: np.ndarray     # [T, nj]
    body_quaternions: np.ndarray    # [T, nbody, 4]

    def as_dict(self) -> Dict[str, np.ndarray]:
        return {k: getattr(self, k) for k in self.__dataclass_fields__}


def _quat_diff_axisangle(q0, q1):
    """axis-angle of q1 * conj(q0)  (preprocessing/transformations.py:83-139 semantics)."""
    d = mjcf.quat_mul(q1, mjcf.quat_conj(q0))
    n = np.linalg.norm(d[1:])
    if n < 1e-12:
        return np.zeros(3)
    ang = 2 * np.arctan2(n, d[0])
    if ang > np.pi:
        ang -= 2 * np.pi
    return d[1:] / n * ang


def synthetic_clip(m: mjcf.Model, free_jnt: bool, T: int = 250, mocap_hz: float = 50.0, seed: int = 0,
                   z_stand: float | None = None, amplitude: float = 0.3, pair_offset=(0.09, 0.063)) -> ReferenceClip:
    """Seeded sinusoidal joint trajectories around qpos0 inside the joint ranges."""
    rng = np.random.default_rng(seed)
    a = m.a
    nq = m.nq
    q0 = a["qpos0"].copy()
    qpos = np.tile(q0, (T, 1))
    t = np.arange(T) / mocap_hz
    for j in range(m.njnt):
        if a["jnt_type"][j] != mjcf.JNT_HINGE:
            continue
        qa = a["jnt_qposadr"][j]
        lo, hi = a["jnt_range"][j]
        f = rng.uniform(0.5, 2.0)
        ph = rng.uniform(0, 2 * np.pi)
        if a["jnt_limited"][j] and hi > lo:
            v = q0[qa] + amplitude * (hi - lo) * 0.5 * np.sin(2 * np.pi * f * t + ph)
            v = np.clip(v, lo, hi)
        else:
            v = q0[qa] + amplitude * np.sin(2 * np.pi * f * t + ph)
        qpos[:, qa] = v
    roots = [j for j in range(m.njnt) if a["jnt_type"][j] == mjcf.JNT_FREE] if free_jnt else []
    for k, j in enumerate(roots):
        qa = a["jnt_qposadr"][j]
        z = q0[qa + 2] if z_stand is None else z_stand
        x, y = 0.1 * t, np.zeros(T)
        yaw = 0.2 * np.sin(2 * np.pi * 0.5 * t)
        if k > 0:
            # every further animal tracks its own copy of the clip: the root trajectory rotated by 180 deg about z and offset so
            # that the animals pass each other at arm's length inside the start-frame range (BASELINE.json configs[3])
            x, y, yaw = pair_offset[0] * k - x, pair_offset[1] * k - y, yaw + np.pi
        qpos[:, qa] = x
        qpos[:, qa + 1] = y
        qpos[:, qa + 2] = z
        qpos[:, qa + 3] = np.cos(yaw / 2)
        qpos[:, qa + 4:qa + 6] = 0.0
        qpos[:, qa + 6] = np.sin(yaw / 2)
        if k > 0:   # ... and the same joint angles as animal 0
            nj = (a["jnt_qposadr"][roots[1]] - 7) if len(roots) > 1 else 0
            qpos[:, qa + 7:qa + 7 + nj] = qpos[:, a["jnt_qposadr"][roots[0]] + 7:a["jnt_qposadr"][roots[0]] + 7 + nj]
    xpos = np.zeros((T, m.nbody, 3))
    xquat = np.zeros((T, m.nbody, 4))
    for k in range(T):
        kin = mjcf.kinematics_np(m, qpos[k])
        xpos[k], xquat[k] = kin["xpos"], kin["xquat"]
    dt = 1.0 / mocap_hz
    if free_jnt:
        # [T, 3 A], [T, 4 A], [T, nq - 7 A]: the roots' columns side by side (A = 1: the reference layout, preprocess.py:23-41)
        qas = [int(a["jnt_qposadr"][j]) for j in roots]
        rootcols = np.concatenate([np.arange(q, q + 7) for q in qas])
        pos = np.concatenate([qpos[:, q:q + 3] for q in qas], 1)
        quat = np.concatenate([qpos[:, q + 3:q + 7] for q in qas], 1)
        joints = np.delete(qpos, rootcols, axis=1)
    else:
        # tethered: preprocess.py:128-129 appends six zero columns; pos/quat terms are unused by the env
        pos, quat, joints = np.zeros((T, 3)), np.tile([1.0, 0, 0, 0], (T, 1)), qpos
    A = max(len(roots), 1)
    vel = np.zeros((T, 3 * A)); ang = np.zeros((T, 3 * A)); jv = np.zeros_like(joints)
    vel[:-1] = (pos[1:] - pos[:-1]) / dt
    jv[:-1] = (joints[1:] - joints[:-1]) / dt
    for k in range(T - 1):
        for r in range(A):
            ang[k, 3 * r:3 * r + 3] = _quat_diff_axisangle(quat[k, 4 * r:4 * r + 4], quat[k + 1, 4 * r:4 * r + 4]) / dt
    jv = np.clip(jv, -20.0, 20.0)  # preprocess.py:131-134
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    return ReferenceClip(position=f32(pos), quaternion=f32(quat), joints=f32(joints), body_positions=f32(xpos),
                         velocity=f32(vel), angular_velocity=f32(ang), joints_velocity=f32(jv), body_quaternions=f32(xquat))

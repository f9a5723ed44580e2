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
    position: np.ndarray            # [T, 3]
    quaternion: np.ndarray          # [T, 4]
    joints: np.ndarray              # [T, nq-7]  (tethered: [T, nq])
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
                   z_stand: float | None = None, amplitude: float = 0.3) -> ReferenceClip:
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
    if free_jnt:
        z = q0[2] if z_stand is None else z_stand
        qpos[:, 0] = 0.1 * t
        qpos[:, 1] = 0.0
        qpos[:, 2] = z
        yaw = 0.2 * np.sin(2 * np.pi * 0.5 * t)
        qpos[:, 3] = np.cos(yaw / 2)
        qpos[:, 4:6] = 0.0
        qpos[:, 6] = np.sin(yaw / 2)
    xpos = np.zeros((T, m.nbody, 3))
    xquat = np.zeros((T, m.nbody, 4))
    for k in range(T):
        kin = mjcf.kinematics_np(m, qpos[k])
        xpos[k], xquat[k] = kin["xpos"], kin["xquat"]
    dt = 1.0 / mocap_hz
    if free_jnt:
        pos, quat, joints = qpos[:, :3], qpos[:, 3:7], qpos[:, 7:]
    else:
        # tethered: preprocess.py:128-129 appends six zero columns; pos/quat terms are unused by the env
        pos, quat, joints = np.zeros((T, 3)), np.tile([1.0, 0, 0, 0], (T, 1)), qpos
    vel = np.zeros((T, 3)); ang = np.zeros((T, 3)); jv = np.zeros_like(joints)
    vel[:-1] = (pos[1:] - pos[:-1]) / dt
    jv[:-1] = (joints[1:] - joints[:-1]) / dt
    for k in range(T - 1):
        ang[k] = _quat_diff_axisangle(quat[k], quat[k + 1]) / dt
    jv = np.clip(jv, -20.0, 20.0)  # preprocess.py:131-134
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    return ReferenceClip(position=f32(pos), quaternion=f32(quat), joints=f32(joints), body_positions=f32(xpos),
                         velocity=f32(vel), angular_velocity=f32(ang), joints_velocity=f32(jv), body_quaternions=f32(xquat))

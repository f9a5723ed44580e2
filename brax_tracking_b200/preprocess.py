"""Mocap joint angles -> ``ReferenceClip``.

Mirrors /root/reference/preprocessing/preprocess.py:99-230: per-frame forward kinematics for the body poses
(``extract_features`` :144-171, here ONE batched launch of the step kernel's tree pass over all T frames instead of a
``lax.scan`` of ``mjx`` kinematics), last-frame padding (:126), six zero columns appended for tethered models (:128-129),
finite-difference velocities with the quaternion-log angular velocity of
``compute_velocity_from_kinematics`` (:207-230; ``transformations.quat_diff`` / ``quat_to_axisangle``) and joint-velocity
clipping (:131-134).  On-disk formats: ``.npz`` with the ``ReferenceClip`` field names and a leading clip axis for multi-clip
files (preprocess.py:254-258) -- written and read here; the reference's own ``.p`` pickles (main.py:57-74: a pickled
``ReferenceClip`` flax dataclass of jax arrays, or a dict of them) -- read here WITHOUT jax / flax through a restricted unpickler
(``load_reference_clip_pickle``).  The reference's HDF5 variant (:233-293) needs h5py, which this image does not have.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np

from . import mjcf
from .clips import ReferenceClip

_TOL = 1e-10


def quat_mul(a, b):
    return np.stack([
        a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3],
        a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2],
        a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1],
        a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0]], -1)


def quat_conj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def quat_diff(source, target):
    """transformations.py:102-114: rotation from source to target = conj(source) * target."""
    return quat_mul(quat_conj(source), target)


def quat_to_axisangle(q):
    """transformations.py:117-139 (batched)."""
    angle = 2 * np.arccos(np.clip(q[..., 0], -1.0, 1.0))
    qn = np.sin(angle / 2)
    wrapped = (angle + np.pi) % (2 * np.pi) - np.pi
    small = angle < _TOL
    axis = q[..., 1:4] / np.where(small, 1.0, qn)[..., None]
    return np.where(small[..., None], 0.0, axis * wrapped[..., None])


def compute_velocity_from_kinematics(qpos_trajectory: np.ndarray, dt: float) -> np.ndarray:
    """preprocess.py:207-230 (free joint in the first 7 columns)."""
    q = np.asarray(qpos_trajectory, dtype=np.float64)
    vt = (q[1:, :3] - q[:-1, :3]) / dt
    d = quat_diff(q[:-1, 3:7], q[1:, 3:7])
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    gyro = quat_to_axisangle(d) / dt
    vj = (q[1:, 7:] - q[:-1, 7:]) / dt
    return np.concatenate([vt, gyro, vj], axis=1)


def process_clip(mocap_qpos: np.ndarray, model: mjcf.Model, kinematics: Optional[Callable] = None, max_qvel: float = 20.0,
                 dt: float = 0.02) -> ReferenceClip:
    """preprocess.py:99-141.  ``kinematics(qpos [T, nq]) -> (xpos [T, nbody, 3], xquat [T, nbody, 4])``: pass
    ``NativeModel.kinematics`` wrapped for numpy (``device_kinematics``) to run the FK on the GPU; defaults to the host
    restatement in ``mjcf.kinematics_np``."""
    q = np.asarray(mocap_qpos, dtype=np.float64)
    T = q.shape[0]
    free = model.a["jnt_type"][0] == mjcf.JNT_FREE
    if kinematics is None:
        ks = [mjcf.kinematics_np(model, q[t]) for t in range(T)]
        xpos, xquat = np.stack([k["xpos"] for k in ks]), np.stack([k["xquat"] for k in ks])
    else:
        xpos, xquat = kinematics(q)
    qn = q.copy()
    if free:
        qn[:, 3:7] /= np.linalg.norm(qn[:, 3:7], axis=1, keepdims=True)  # kinematics stores the normalised quaternion
        position, quaternion, joints = qn[:, :3], qn[:, 3:7], qn[:, 7:]
    else:
        position, quaternion, joints = np.zeros((T, 3)), np.tile([1.0, 0, 0, 0], (T, 1)), qn
    padded = np.concatenate([q, q[-1:]], axis=0)
    if not free:
        padded = np.concatenate([padded, np.zeros((padded.shape[0], 6))], axis=1)   # preprocess.py:128-129 (appended, sic)
    qvel = compute_velocity_from_kinematics(padded, dt)
    qvel[:, 6:] = np.clip(qvel[:, 6:], -max_qvel, max_qvel)
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    return ReferenceClip(position=f32(position), quaternion=f32(quaternion), joints=f32(joints), body_positions=f32(xpos),
                         velocity=f32(qvel[:, :3]), angular_velocity=f32(qvel[:, 3:6]), joints_velocity=f32(qvel[:, 6:]),
                         body_quaternions=f32(xquat))


def device_kinematics(native_model) -> Callable:
    """numpy-in / numpy-out wrapper of ``NativeModel.kinematics`` (one launch for the whole clip)."""
    import torch

    def f(q):
        t = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(native_model._dev())
        xp, xq = native_model.kinematics(t)
        return xp.cpu().numpy().astype(np.float64), xq.cpu().numpy().astype(np.float64)
    return f


def save_reference_clip(path: str, clips: Dict[str, ReferenceClip] | ReferenceClip) -> None:
    """One clip, or {name: clip} stacked on a leading clip axis (preprocess.py:233-258)."""
    if isinstance(clips, ReferenceClip):
        np.savez_compressed(path, **clips.as_dict())
        return
    names = sorted(clips)
    out = {k: np.stack([getattr(clips[n], k) for n in names]) for k in ReferenceClip.__dataclass_fields__}
    np.savez_compressed(path, __clip_names__=np.array(names), **out)


def load_reference_clip(path: str, clip_idx: Optional[int] = None) -> ReferenceClip:
    z = np.load(path)
    if "__clip_names__" in z.files:
        if clip_idx is None:
            raise ValueError("multi-clip file: pass clip_idx")
        return ReferenceClip(**{k: z[k][clip_idx] for k in ReferenceClip.__dataclass_fields__})
    return ReferenceClip(**{k: z[k] for k in ReferenceClip.__dataclass_fields__})


# ---------------------------------------------------------------------------------------------- the reference's .p pickles
class _PickledStruct:
    """Stand-in for any class the pickle names that is not importable here (``preprocessing.preprocess.ReferenceClip``, a
    flax.struct dataclass): keeps the instance ``__dict__`` / state."""

    def __init__(self, *a, **k):
        self.__dict__.update(k)

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


def _reconstruct_jax_array(fun, args, arr_state, aval_state=None):
    """``jax._src.array._reconstruct_array``: a pickled ``jax.Array`` is (numpy reconstructor, its args, the ndarray state, the aval
    state) -- rebuild the numpy array and drop the device placement."""
    arr = fun(*args)
    arr.__setstate__(arr_state)
    return np.asarray(arr)


def load_reference_clip_pickle(path: str):
    """Reads what /root/reference/main.py:57-74 reads with ``pickle.load``: a ``ReferenceClip`` instance (or a dict / list of them)
    whose leaves are jax or numpy arrays.  Only numpy reconstruction and the array / dataclass stand-ins above are allowed to run
    (a pickle is code: nothing else is ever imported or called).  Returns a ``ReferenceClip``, or {name: ReferenceClip} / a list."""
    import pickle

    allowed = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
               ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("collections", "OrderedDict")}

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if (module, name) in allowed:
                return super().find_class(module, name)
            if name == "_reconstruct_array" and module.startswith("jax"):
                return _reconstruct_jax_array
            if module.startswith(("jax", "jaxlib", "flax")) and name in ("ShapedArray", "UnspecifiedValue", "ArrayImpl"):
                return _PickledStruct
            if name == "ReferenceClip" or module.startswith("preprocessing"):
                return _PickledStruct
            raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a clip pickle")

    with open(path, "rb") as f:
        obj = Unpickler(f).load()

    def to_clip(o):
        d = o if isinstance(o, dict) else o.__dict__
        missing = [k for k in ReferenceClip.__dataclass_fields__ if k not in d]
        if missing:
            raise ValueError(f"not a ReferenceClip: missing {missing}")
        return ReferenceClip(**{k: np.ascontiguousarray(np.asarray(d[k]), dtype=np.float32) for k in ReferenceClip.__dataclass_fields__})

    if isinstance(obj, (list, tuple)):
        return [to_clip(o) for o in obj]
    if isinstance(obj, dict) and not set(ReferenceClip.__dataclass_fields__) <= set(obj):
        return {k: to_clip(v) for k, v in obj.items()}
    return to_clip(obj)

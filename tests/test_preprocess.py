"""Clip preprocessing and the render-rollout reset (rank 3)."""
import numpy as np
import pytest

import common
from backends import EmuBackend
from brax_tracking_b200 import mjcf, preprocess


def test_quaternion_log_velocity_is_exact_for_constant_rotation():
    dt, w = 0.02, np.array([0.3, -1.1, 2.0])
    T = 20
    ang = np.linalg.norm(w) * dt * np.arange(T)
    ax = w / np.linalg.norm(w)
    quat = np.concatenate([np.cos(ang / 2)[:, None], np.sin(ang / 2)[:, None] * ax[None]], 1)
    q = np.concatenate([np.arange(T)[:, None] * dt * np.array([[1.0, 0.5, -0.2]]), quat, np.linspace(0, 1, T)[:, None]], 1)
    v = preprocess.compute_velocity_from_kinematics(q, dt)
    np.testing.assert_allclose(v[:, :3], np.tile([1.0, 0.5, -0.2], (T - 1, 1)), atol=1e-9)
    np.testing.assert_allclose(v[:, 3:6], np.tile(w, (T - 1, 1)), atol=1e-7)      # body-frame rotation about a fixed axis
    np.testing.assert_allclose(v[:, 6], 1.0 / (T - 1) / dt, atol=1e-9)


@pytest.mark.parametrize("name", ["rodent", "fly_tethered"])
def test_process_clip_shapes_and_kernel_fk(name, tmp_path):
    m, cfg, clip, tables = common.setup(name)
    rng = np.random.default_rng(0)
    T = 12
    q = np.tile(m.qpos0, (T, 1)) + 0.05 * rng.standard_normal((T, m.nq))
    c_host = preprocess.process_clip(q, m)
    # the same FK through the kernels' tree pass (host emulation here; NativeModel.kinematics on the GPU)
    b = EmuBackend(tables)

    def emu_fk(qq):
        st = b.e.new_state(qq.shape[0]); st["qpos"][:] = qq
        sc, _, _ = b.forward_debug(st, None, stop=1)
        return (common.region(tables, sc, "xpos", 3 * m.nbody).reshape(-1, m.nbody, 3).astype(np.float64),
                common.region(tables, sc, "xquat", 4 * m.nbody).reshape(-1, m.nbody, 4).astype(np.float64))
    c_dev = preprocess.process_clip(q, m, kinematics=emu_fk)
    np.testing.assert_allclose(c_dev.body_positions, c_host.body_positions, atol=2e-6)
    free = cfg["free_jnt"]
    assert c_host.joints.shape == (T, m.nq - 7 if free else m.nq) and c_host.body_positions.shape == (T, m.nbody, 3)
    assert c_host.velocity.shape == (T, 3) and c_host.joints_velocity.shape[0] == T           # padded last frame
    assert np.abs(c_host.joints_velocity).max() <= 20.0 and not c_host.joints_velocity[-1].any()
    p = str(tmp_path / "clip.npz")
    preprocess.save_reference_clip(p, c_host)
    back = preprocess.load_reference_clip(p)
    assert all(np.array_equal(getattr(back, k), getattr(c_host, k)) for k in c_host.as_dict())
    preprocess.save_reference_clip(str(tmp_path / "multi.npz"), {"a": c_host, "b": c_dev})
    assert np.array_equal(preprocess.load_reference_clip(str(tmp_path / "multi.npz"), 1).joints, c_dev.joints)


def test_render_rollout_reset_starts_at_frame_zero():
    """RenderRolloutWrapperTracking.reset (custom_wrappers.py:85-125): frame 0, split(rng, 3), qpos0 + noise."""
    m, cfg, clip, tables = common.setup("rodent")
    _, eo = common.oracles("rodent")
    b = EmuBackend(tables)
    keys = common.jax_keys(8, seed=4)
    st, out = b.reset(keys, fixed_start_frame=0)
    s0 = eo.reset(keys, fixed_start_frame=0)
    assert not out["info_i"].any()
    assert np.array_equal(st["qvel"], s0["pipeline_state"]["qvel"])
    np.testing.assert_allclose(st["qpos"], s0["pipeline_state"]["qpos"], atol=1.2e-7)
    np.testing.assert_allclose(out["obs"], s0["obs"], atol=2e-5)
    st_t, _ = b.reset(keys)                                    # the training reset draws different noise (split(rng, 4))
    assert not np.array_equal(st_t["qvel"], st["qvel"])
    assert np.abs(st["qpos"][:, :2]).max() < 2e-3              # not seeded from the clip


def test_reference_pickle_reader_without_jax(tmp_path):
    """main.py:57-74 unpickles a `preprocessing.preprocess.ReferenceClip` (flax dataclass of jax arrays).  Neither module exists
    here: the restricted unpickler maps the class to a stand-in, rebuilds `jax.Array` leaves from their numpy payload
    (`jax._src.array._reconstruct_array`) and refuses everything else."""
    import pickle
    import sys
    import types
    m, cfg, clip, _ = common.setup("rodent")
    fields = {k: np.asarray(v)[:7] for k, v in clip.items()}
    # build the pickle the way the reference would: fake module path for the class, jax-style reduce for two of the arrays
    pre = types.ModuleType("preprocessing"); pp = types.ModuleType("preprocessing.preprocess")
    jx = types.ModuleType("jax"); jsrc = types.ModuleType("jax._src"); jarr = types.ModuleType("jax._src.array")

    class ReferenceClip:                                    # pickled by reference: module path + __dict__
        pass
    ReferenceClip.__module__, ReferenceClip.__qualname__ = "preprocessing.preprocess", "ReferenceClip"
    pp.ReferenceClip = ReferenceClip

    def _reconstruct_array(fun, args, arr_state, aval_state):
        raise AssertionError("only referenced by name")
    _reconstruct_array.__module__, _reconstruct_array.__qualname__ = "jax._src.array", "_reconstruct_array"
    jarr._reconstruct_array = _reconstruct_array

    class FakeJaxArray:
        def __init__(self, a):
            self.a = np.ascontiguousarray(a)

        def __reduce__(self):
            fun, args, state = self.a.__reduce__()
            return _reconstruct_array, (fun, args, state, ("aval",))
    mods = {"preprocessing": pre, "preprocessing.preprocess": pp, "jax": jx, "jax._src": jsrc, "jax._src.array": jarr}
    sys.modules.update(mods)
    try:
        obj = ReferenceClip()
        for k, v in fields.items():
            setattr(obj, k, FakeJaxArray(v) if k in ("position", "joints") else v)
        p = str(tmp_path / "clip.p")
        with open(p, "wb") as f:
            pickle.dump(obj, f)
        with open(str(tmp_path / "multi.p"), "wb") as f:
            pickle.dump({"walk": obj, "rear": obj}, f)
    finally:
        for k in mods:
            sys.modules.pop(k, None)
    back = preprocess.load_reference_clip_pickle(p)
    for k, v in fields.items():
        assert np.array_equal(getattr(back, k), v.astype(np.float32)), k
    multi = preprocess.load_reference_clip_pickle(str(tmp_path / "multi.p"))
    assert sorted(multi) == ["rear", "walk"] and np.array_equal(multi["walk"].joints, back.joints)
    with open(str(tmp_path / "evil.p"), "wb") as f:          # a pickle is code: anything outside the whitelist is refused
        pickle.dump(preprocess.process_clip, f)
    with pytest.raises(pickle.UnpicklingError):
        preprocess.load_reference_clip_pickle(str(tmp_path / "evil.p"))

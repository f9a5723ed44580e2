"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

ctypes front-end of ``mjx_oracle.c`` (the CPU restatement of ``mjx.forward`` / ``mjx.step``
reached through /root/reference/envs/fruitfly.py:500).  **parity unpinned**: the reference holds
no golden vectors for this path and MJX is not importable here.

Use ``Oracle(model, dtype)`` for a single environment with every intermediate exposed, and
``Oracle.step_batch`` for (threaded) batches -- the latter is what the CPU baseline times.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_M_INT = ["nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "ntendon", "npair", "ncon", "nefc_max", "nlimit",
          "cone", "iterations", "ls_iterations"]
_M_DBL = ["timestep", ("gravity", 3), "density", "viscosity", "impratio", "tolerance", "ls_tolerance", "meaninertia"]
_M_IPTR = ["body_parentid", "body_rootid", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum",
           "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited",
           "dof_bodyid", "dof_jntid", "dof_parentid",
           "geom_type", "geom_bodyid",
           "pair_geom", "pair_condim", "pair_ncon",
           "tendon_adr", "tendon_num", "wrap_jntid",
           "actuator_trntype", "actuator_trnid", "actuator_dyntype", "actuator_gaintype", "actuator_biastype",
           "actuator_ctrllimited", "actuator_forcelimited", "actuator_actadr"]
_M_DPTR = ["body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_invweight0",
           "jnt_pos", "jnt_axis", "jnt_range", "jnt_stiffness", "jnt_solref", "jnt_solimp", "jnt_margin",
           "qpos0", "qpos_spring", "dof_armature", "dof_damping", "dof_invweight0",
           "geom_pos", "geom_quat", "geom_size",
           "pair_friction", "pair_solref", "pair_solimp", "pair_margin", "pair_gap",
           "wrap_coef",
           "actuator_gear", "actuator_gainprm", "actuator_biasprm", "actuator_dynprm", "actuator_ctrlrange",
           "actuator_forcerange"]


class _OModel(C.Structure):
    _fields_ = ([(n, C.c_int) for n in _M_INT]
                + [((n if isinstance(n, str) else n[0]), (C.c_double if isinstance(n, str) else C.c_double * n[1])) for n in _M_DBL]
                + [(n, C.POINTER(C.c_int)) for n in _M_IPTR]
                + [(n, C.POINTER(C.c_double)) for n in _M_DPTR])


def _data_fields(m):
    nq, nv, nu, na, nb, nj, ng = m.nq, m.nv, m.nu, m.na, m.nbody, m.njnt, m.ngeom
    nc, ne = m._ncon, m._nefc_max
    # (name, size, is_int)  -- order must match struct OData
    return [
        ("qpos", nq, 0), ("qvel", nv, 0), ("act", max(na, 1), 0), ("ctrl", max(nu, 1), 0), ("qacc_warmstart", nv, 0), ("time", 1, 0),
        ("xpos", nb * 3, 0), ("xquat", nb * 4, 0), ("xmat", nb * 9, 0), ("xipos", nb * 3, 0), ("ximat", nb * 9, 0),
        ("xanchor", nj * 3, 0), ("xaxis", nj * 3, 0), ("geom_xpos", ng * 3, 0), ("geom_xmat", ng * 9, 0),
        ("subtree_com", nb * 3, 0), ("cinert", nb * 10, 0), ("cdof", nv * 6, 0), ("crb", nb * 10, 0), ("qM", nv * nv, 0), ("qLD", nv * nv, 0),
        ("cvel", nb * 6, 0), ("cdof_dot", nv * 6, 0), ("qfrc_bias", nv, 0), ("qfrc_passive", nv, 0),
        ("actuator_length", max(nu, 1), 0), ("actuator_velocity", max(nu, 1), 0), ("actuator_force", max(nu, 1), 0),
        ("act_dot", max(na, 1), 0), ("qfrc_actuator", nv, 0), ("actuator_moment", max(nu * nv, 1), 0),
        ("qfrc_smooth", nv, 0), ("qacc_smooth", nv, 0), ("qacc", nv, 0), ("qfrc_constraint", nv, 0),
        ("con_dist", max(nc, 1), 0), ("con_pos", max(nc, 1) * 3, 0), ("con_frame", max(nc, 1) * 9, 0), ("con_friction", max(nc, 1) * 5, 0),
        ("con_solref", max(nc, 1) * 2, 0), ("con_solimp", max(nc, 1) * 5, 0), ("con_includemargin", max(nc, 1), 0),
        ("con_geom", max(nc, 1) * 2, 1), ("con_dim", max(nc, 1), 1),
        ("efc_J", max(ne, 1) * nv, 0), ("efc_D", max(ne, 1), 0), ("efc_aref", max(ne, 1), 0), ("efc_pos", max(ne, 1), 0), ("efc_force", max(ne, 1), 0),
        ("efc_type", max(ne, 1), 1), ("efc_id", max(ne, 1), 1),
        ("nefc", 1, 1), ("solver_niter", 1, 1),
        ("scratch", 40 * nv + 16 * max(ne, 1) + 3 * nv * nv + 16 * nb + 64, 0),
    ]


def build(force=False):
    """Compile the two oracle libraries with the committed Makefile."""
    out = os.path.join(_HERE, "_build")
    need = force or not all(os.path.exists(os.path.join(out, f)) for f in ("liboracle_f64.so", "liboracle_f32.so"))
    if not need:
        src_t = os.path.getmtime(os.path.join(_HERE, "mjx_oracle.c"))
        need = any(os.path.getmtime(os.path.join(out, f)) < src_t for f in ("liboracle_f64.so", "liboracle_f32.so"))
    if need:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


_LIBS = {}


def _lib(dtype):
    key = np.dtype(dtype).name
    if key not in _LIBS:
        build()
        name = "liboracle_f64.so" if key == "float64" else "liboracle_f32.so"
        lib = C.CDLL(os.path.join(_HERE, "_build", name))
        assert lib.o_sizeof_real() == np.dtype(dtype).itemsize
        _LIBS[key] = lib
    return _LIBS[key]


class OracleModel:
    """Holds the C model struct (and keeps the numpy arrays it points to alive)."""

    def __init__(self, model):
        self.model = model
        a = model.a
        self._keep = {}
        s = _OModel()
        nlimit = int(np.sum((a["jnt_limited"] != 0) & (a["jnt_type"] == 3)))
        ncon = int(a["pair_ncon"].sum())
        rows = 0
        for p in range(len(a["pair_ncon"])):
            dim = int(a["pair_condim"][p])
            per = 1 if dim == 1 else (2 * (dim - 1) if model.cone == 0 else dim)
            rows += per * int(a["pair_ncon"][p])
        self.ncon, self.nefc_max = ncon, nlimit + rows
        vals = dict(nq=model.nq, nv=model.nv, nu=model.nu, na=model.na, nbody=model.nbody, njnt=model.njnt,
                    ngeom=model.ngeom, ntendon=model.ntendon, npair=len(a["pair_ncon"]), ncon=ncon,
                    nefc_max=self.nefc_max, nlimit=nlimit, cone=model.cone, iterations=model.iterations,
                    ls_iterations=model.ls_iterations)
        for n in _M_INT:
            setattr(s, n, int(vals[n]))
        for n in _M_DBL:
            if isinstance(n, str):
                setattr(s, n, float(getattr(model, n)))
            else:
                arr = (C.c_double * n[1])(*[float(x) for x in getattr(model, n[0])])
                setattr(s, n[0], arr)
        for n in _M_IPTR:
            arr = np.ascontiguousarray(a[n], dtype=np.int32).ravel()
            if arr.size == 0:
                arr = np.zeros(1, dtype=np.int32)
            self._keep[n] = arr
            setattr(s, n, arr.ctypes.data_as(C.POINTER(C.c_int)))
        for n in _M_DPTR:
            arr = np.ascontiguousarray(a[n], dtype=np.float64).ravel()
            if arr.size == 0:
                arr = np.zeros(1, dtype=np.float64)
            self._keep[n] = arr
            setattr(s, n, arr.ctypes.data_as(C.POINTER(C.c_double)))
        self.struct = s
        # sizes used by _data_fields
        self.nq, self.nv, self.nu, self.na = model.nq, model.nv, model.nu, model.na
        self.nbody, self.njnt, self.ngeom = model.nbody, model.njnt, model.ngeom
        self._ncon, self._nefc_max = self.ncon, self.nefc_max


class OracleData:
    def __init__(self, om: OracleModel, dtype):
        self.dtype = np.dtype(dtype)
        fields = _data_fields(om)
        rtype = C.c_double if self.dtype == np.float64 else C.c_float
        cls = type("_OData", (C.Structure,), {"_fields_": [(n, C.POINTER(C.c_int if isint else rtype)) for n, _, isint in fields]})
        self.struct = cls()
        self.arr = {}
        for n, size, isint in fields:
            arr = np.zeros(size, dtype=np.int32 if isint else self.dtype)
            self.arr[n] = arr
            setattr(self.struct, n, arr.ctypes.data_as(C.POINTER(C.c_int if isint else rtype)))
        self.arr["xquat"].reshape(-1, 4)[:, 0] = 1

    def __getattr__(self, k):
        arr = self.__dict__.get("arr")
        if arr is not None and k in arr:
            return arr[k]
        raise AttributeError(k)


class Oracle:
    """Single-environment oracle with all intermediates, plus batched stepping."""

    def __init__(self, model, dtype=np.float64):
        self.m = model
        self.om = OracleModel(model)
        self.dtype = np.dtype(dtype)
        self.lib = _lib(dtype)
        self.d = OracleData(self.om, dtype)
        self._pool_data = []

    # -- single env ------------------------------------------------------------------------
    def set_state(self, qpos, qvel, act=None, warmstart=None, ctrl=None, time=0.0):
        d = self.d
        d.qpos[:] = qpos
        d.qvel[:] = qvel
        d.act[:] = 0
        if act is not None and self.m.na:
            d.act[: self.m.na] = act
        d.qacc_warmstart[:] = 0 if warmstart is None else warmstart
        d.ctrl[:] = 0
        if ctrl is not None:
            d.ctrl[: self.m.nu] = ctrl
        d.time[0] = time

    def call(self, name, d=None):
        d = d or self.d
        getattr(self.lib, name)(C.byref(self.om.struct), C.byref(d.struct))

    def forward(self):
        self.call("o_forward")

    def step(self):
        self.call("o_step")

    # -- batched (threads; ctypes releases the GIL) ------------------------------------------
    def _get_datas(self, n):
        while len(self._pool_data) < n:
            self._pool_data.append(OracleData(self.om, self.dtype))
        return self._pool_data[:n]

    def pipeline_batch(self, state, ctrl, n_frames, threads=None, forward_only=False):
        """state: dict of [N, dim] arrays (qpos, qvel, act, qacc_warmstart, time); ctrl [N, nu].
        Runs n_frames x mjx.step per env (brax PipelineEnv.pipeline_step) or one mjx.forward (pipeline_init).
        Returns the new state dict (+ xpos from the last forward, as MJX leaves it in Data)."""
        N = state["qpos"].shape[0]
        threads = threads or min(os.cpu_count() or 1, N)
        datas = self._get_datas(threads)
        out = {k: np.array(v, dtype=self.dtype, copy=True) for k, v in state.items() if k in ("qpos", "qvel", "act", "qacc_warmstart", "time")}
        out["xpos"] = np.zeros((N, self.m.nbody, 3), dtype=self.dtype)
        out["qacc"] = np.zeros((N, self.m.nv), dtype=self.dtype)
        na, nu = self.m.na, self.m.nu

        def work(t):
            d = datas[t]
            for e in range(t, N, threads):
                d.qpos[:] = out["qpos"][e]
                d.qvel[:] = out["qvel"][e]
                if na:
                    d.act[:na] = out["act"][e]
                d.qacc_warmstart[:] = out["qacc_warmstart"][e]
                d.time[0] = out["time"][e]
                if nu:   # pipeline_init runs mjx.forward on a fresh mjx.make_data: ctrl = 0 (the pooled buffers keep the last call's)
                    d.ctrl[:nu] = ctrl[e] if ctrl is not None else 0
                if forward_only:
                    self.lib.o_forward(C.byref(self.om.struct), C.byref(d.struct))
                else:
                    for _ in range(n_frames):
                        self.lib.o_step(C.byref(self.om.struct), C.byref(d.struct))
                out["qpos"][e] = d.qpos
                out["qvel"][e] = d.qvel
                if na:
                    out["act"][e] = d.act[:na]
                out["qacc_warmstart"][e] = d.qacc_warmstart
                out["time"][e] = d.time[0]
                out["xpos"][e] = d.xpos.reshape(-1, 3)
                out["qacc"][e] = d.qacc

        if threads == 1:
            work(0)
        else:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(work, range(threads)))
        return out

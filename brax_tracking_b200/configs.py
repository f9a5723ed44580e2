"""Environment constants of the reference's dataset configs, and name -> id resolution.

Values are those of /root/reference/configs/dataset/env_config.yaml:7-93 (rodent),
configs/dataset/fly.yaml:8-149 (tethered fly) and fly_freejnt.yaml (free-root fly; differs only in
``free_jnt``), plus the constructor defaults of /root/reference/envs/fruitfly.py:346-378 for everything the
YAMLs do not override (``ref_len=5``, ``mocap_hz=50``, ``reset_noise_scale=1e-3``).  ``episode_length`` follows
/root/reference/main.py:86.  Name lists keep the YAML typos: ``mj_name2id`` returns -1 for them and the env
indexes with -1, which ``model.pack`` reproduces with JAX gather semantics.
"""
from __future__ import annotations

from typing import Dict, List

from . import mjcf

_COMMON = dict(
    mocap_hz=50, ref_len=5, reset_noise_scale=1e-3, start_frame_range=44,  # fruitfly.py:354,358,375,453
    iterations=4, ls_iterations=4, physics_steps_per_control_step=5,
    bad_pose_dist=1000.0, bad_quat_dist=1000.0, ctrl_cost_weight=0.01, quat_reward_weight=1.0,
    angvel_reward_weight=0.0, bodypos_reward_weight=1.0, endeff_reward_weight=1.0, healthy_reward=0.25,
    terminate_when_unhealthy=True, clip_length=250, ref_traj_length=5,
)

RODENT_ENV_ARGS = dict(
    _COMMON,
    scale_factor=0.9, too_far_dist=0.01, pos_reward_weight=1.0, joint_reward_weight=1.0,
    healthy_z_range=(0.0325, 0.5), free_jnt=True, seed_root_from_clip=True,  # rodent.py:154-159
    center_of_mass="torso",
    # `_endeff_idxs` come from appendage_names in rodent.py:113-119
    end_eff_names=["foot_L", "foot_R", "hand_L", "hand_R", "skull"],
    body_names=["torso", "pelvis", "upper_leg_L", "lower_leg_L", "foot_L", "upper_leg_R", "lower_leg_R", "foot_R", "skull",
                "jaw", "scapula_L", "upper_arm_L", "lower_arm_L", "finger_L", "scapula_R", "upper_arm_R", "lower_arm_R",
                "finger_R"],
    joint_names=["vertebra_1_extend", "hip_L_supinate", "hip_L_abduct", "hip_L_extend", "knee_L", "ankle_L", "toe_L",
                 "hip_R_supinate", "hip_R_abduct", "hip_R_extend", "knee_R", "ankle_R", "toe_R", "vertebra_C11_extend",
                 "vertebra_cervical_1_bend", "vertebra_axis_twist", "atlas", "mandible", "scapula_L_supinate",
                 "scapula_L_abduct", "scapula_L_extend", "shoulder_L", "shoulder_sup_L", "elbow_L", "wrist_L",
                 "scapula_R_supinate", "scapula_R_abduct", "scapula_R_extend", "shoulder_R", "shoulder_sup_R", "elbow_R",
                 "wrist_R", "finger_R"],
)


# BASELINE.json configs[3]: assets/rodent_pair.xml (two replicas of the animal in one world) with the rodent env semantics per
# animal; SURVEY.md Appendix C.3 variant (i), DESIGN.md "config 4".  `inter_animal_pairs`: the opened capsule-capsule pairs
# (geom names without the `_collision-k` suffix; animal 0's geom first) -- declared deviation: the file as committed lets the
# animals collide with the floor only (rodent_pair.xml:31-40).
_FORE, _HIND = ("humerus_L", "humerus_R"), ("upper_leg_L0", "upper_leg_R0")
RODENT_PAIR_ENV_ARGS = dict(
    RODENT_ENV_ARGS, animal_suffixes=["-0", "-1"],
    inter_animal_pairs=([(a, b) for a in _FORE for b in _FORE] + [(a, b) for a in _HIND for b in _HIND]
                        + [(_FORE[i], _HIND[i]) for i in range(2)] + [(_HIND[i], _FORE[i]) for i in range(2)]))


def pair_geom_names(env_args):
    return [(f"{a}_collision-0", f"{b}_collision-1") for a, b in env_args["inter_animal_pairs"]]


def _fly_names():
    bodies = ["thorax", "head", "rostrum", "haustellum", "labrum_left", "labrum_right", "antenna_left", "antenna_right",
              "wing_left", "wing_right", "abdomen"] + [f"abdomen_{k}" for k in range(2, 8)] + ["haltere_left", "haltere_right"]
    joints: List[str] = []
    for T in ("T1", "T2", "T3"):
        for side in ("left", "right"):
            bodies += [f"{seg}_{T}_{side}" for seg in ("coxa", "femur", "tibia", "tarsus", "tarsus2", "tarsus3", "tarsus4", "claw")]
            if side == "left":
                joints += [f"coxa_flexion_{T}_left", f"coxa_twist_{T}_left", f"femur_{T}_left", f"femur_twist_{T}_left",
                           f"tibia_{T}_left", f"tarsus_{T}_left"]
            else:
                # typos as committed in fly.yaml:120-122,132-136,144-148
                last = {"T1": "tarsus_T1_right", "T2": "tarsus_T2_righ", "T3": "tarsus_T3_rig"}[T]
                joints += [f"coxa_flexion_{T}_right", f"oxa_twist_{T}_right", f"emur_{T}_right", f"emur_twist_{T}_right",
                           f"tibia_{T}_right", last]
    return bodies, joints


_FLY_BODIES, _FLY_JOINTS = _fly_names()

FLY_ENV_ARGS = dict(
    _COMMON,
    scale_factor=1, too_far_dist=0.1, pos_reward_weight=0.0, joint_reward_weight=50.0,
    healthy_z_range=(-0.05, 0.1), free_jnt=False, seed_root_from_clip=False, center_of_mass="thorax",
    end_eff_names=[f"claw_{T}_{s}" for T in ("T1", "T2", "T3") for s in ("left", "right")],
    body_names=_FLY_BODIES, joint_names=_FLY_JOINTS,
)
FLY_FREEJNT_ENV_ARGS = dict(FLY_ENV_ARGS, free_jnt=True)


def resolve(m: mjcf.Model, env_args: dict) -> Dict:
    """Env kwargs -> the constant dict consumed by ``model.pack`` (what the env constructor derives,
    fruitfly.py:405-447)."""
    max_sub = int(1.0 / (env_args["mocap_hz"] * m.timestep))
    n_frames = env_args["physics_steps_per_control_step"]
    if max_sub % n_frames != 0:
        raise ValueError(f"physics_steps_per_control_step ({n_frames}) must be a factor of ({max_sub})")  # fruitfly.py:414-415
    steps_for_cur_frame = max_sub / n_frames
    cfg = {k: env_args[k] for k in (
        "ref_len", "reset_noise_scale", "start_frame_range", "too_far_dist", "bad_pose_dist", "bad_quat_dist",
        "ctrl_cost_weight", "pos_reward_weight", "quat_reward_weight", "joint_reward_weight", "angvel_reward_weight",
        "bodypos_reward_weight", "endeff_reward_weight", "healthy_reward", "healthy_z_range", "terminate_when_unhealthy",
        "free_jnt", "seed_root_from_clip")}
    cfg["n_frames"] = n_frames
    cfg["steps_for_cur_frame"] = steps_for_cur_frame
    # main.py:86
    cfg["episode_length"] = int((env_args["clip_length"] - 50 - env_args["ref_traj_length"]) * steps_for_cur_frame)
    # one record per tracked animal = per free root (the tethered fly: one pseudo-animal spanning all of qpos).  Every reference
    # env has exactly one; the two-rodent model of BASELINE.json configs[3] has two, whose names carry the replicate suffixes
    # of assets/rodent_pair.xml:163 (`-0`, `-1`).  Joint ids are taken relative to the animal's root joint, so that each animal
    # reproduces the single-animal indexing, quirks included (ids used as columns of `qpos[7:]`).
    roots = [j for j in range(m.njnt) if m.jnt_type[j] == mjcf.JNT_FREE] if env_args["free_jnt"] else []
    suffixes = env_args.get("animal_suffixes") or [""]
    if roots and len(suffixes) != len(roots):
        raise ValueError(f"the model has {len(roots)} free roots but the env names {len(suffixes)} animals")
    animals, jbase = [], 0
    for k, sfx in enumerate(suffixes):
        if roots:
            j0 = roots[k]
            j1 = roots[k + 1] if k + 1 < len(roots) else m.njnt
            qadr, dadr, nj = int(m.jnt_qposadr[j0]), int(m.jnt_dofadr[j0]), j1 - j0 - 1
            if any(m.jnt_type[j] != mjcf.JNT_HINGE for j in range(j0 + 1, j1)):
                raise NotImplementedError("animals are a free root followed by hinges")
        else:
            j0, qadr, dadr, nj = 0, 0, 0, m.nq

        def jid(n):
            i = m.name2id("joint", n + sfx)
            return i - j0 if i >= 0 else -1
        animals.append(dict(
            qadr=qadr, dadr=dadr, nj=nj, jbase=jbase, torso_idx=m.name2id("body", env_args["center_of_mass"] + sfx),
            joint_idxs=[jid(n) for n in env_args["joint_names"]],
            body_idxs=[m.name2id("body", n + sfx) for n in env_args["body_names"]],
            endeff_idxs=[m.name2id("body", n + sfx) for n in env_args["end_eff_names"]]))
        jbase += nj
    cfg["animals"] = animals
    for k in ("torso_idx", "joint_idxs", "body_idxs", "endeff_idxs"):   # the env attributes other layers read: animal 0's
        cfg[k] = animals[0][k]
    return cfg

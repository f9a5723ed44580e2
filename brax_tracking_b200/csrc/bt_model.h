// Device-side model description for the fused physics + tracking-reward step.
//
// All model constants are uploaded once by bt_model_create() as *named* tables: the X-macro lists below
// name exactly the arrays produced by the host packer (brax_tracking_b200/model.py::pack).  Kernels take the
// struct by value (__grid_constant__) and read the tables through the read-only path; per-environment state
// lives in shared memory for the duration of a control step ("data layout").
#pragma once
#include <stdint.h>

// ---- scalar ints (1-element int tables, by name) ----
#define BT_INT_SCALARS(X) \
  X(nq) X(nv) X(nu) X(na) X(nbody) X(njnt) X(nhpass) X(napass) X(nbanc) X(ncon) X(ncb) X(nwgrp) X(nmerge) X(nchain) \
  X(cone) X(iterations) X(ls_iterations) X(n_frames) X(sync_mode) X(ncross) X(poison) X(jt_seg_steps) X(wgrp_contig)                                          \
  /* env layer */                                                                                               \
  X(free_jnt) X(seed_root_from_clip) X(ref_len) X(clip_len) X(clip_nj) X(n_joint_idxs) X(n_body_idxs) X(n_animals) X(n_clips) \
  X(n_endeff_idxs) X(torso_idx) X(terminate_when_unhealthy) X(steps_for_cur_frame) X(episode_length)            \
  X(start_frame_range) X(obs_size)                                                                              \
  /* per-environment scratch layout (offsets in floats) */                                                      \
  X(o_qpos) X(o_qvel) X(o_act) X(o_ctrl) X(o_warm) X(o_xpos) X(o_xquat) X(o_cdof) X(o_crb) X(o_Dinv) X(o_Dd)    \
  X(o_cbJ) X(o_pvec) X(o_T) X(o_ref) X(o_aforce) X(o_actdot) X(o_qfrc_smooth) X(o_qacc_smooth) X(o_qacc) X(o_x) \
  X(o_search) X(o_qfrc_c) X(o_tmpv) X(o_wrench) X(o_cbA) X(smem_floats)                                            \
  /* CTA-shared constant records (offsets into sh_tab; the first sh_stage_floats floats are staged in shared memory) */ \
  X(sh_stage_floats) X(sho_body_rec) X(sho_bl_rec) X(sho_jnt_rec) X(sho_wrap_rec) X(sho_dofact_rec)

// ---- scalar floats ----
#define BT_FLT_SCALARS(X) \
  X(timestep) X(grav_x) X(grav_y) X(grav_z) X(density) X(viscosity) X(impratio) X(tolerance) X(ls_tolerance)    \
  X(meaninertia)                                                                                                \
  X(too_far_dist) X(bad_pose_dist) X(bad_quat_dist) X(ctrl_cost_weight) X(pos_reward_weight)                    \
  X(quat_reward_weight) X(joint_reward_weight) X(angvel_reward_weight) X(bodypos_reward_weight)                 \
  X(endeff_reward_weight) X(healthy_reward) X(healthy_z_min) X(healthy_z_max) X(reset_noise_scale)

// ---- int tables ----
// (model.py::pack emits more tables than the kernels bind: the unpacked per-field arrays behind the packed records below
// stay in the dict for the host-side consumers -- tests, the oracle adapters, tools)
#define BT_INT_TABLES(X) \
  X(body_parentid) X(body_ref) X(cmp_adr) X(cmp_item)                                                           \
  X(jnt_type) X(jnt_qposadr) X(jnt_dofadr) X(jnt_bodyid)                                                        \
  X(dof_qposadr) X(dof_limited) X(dof_vflag) X(dof_irec) X(merge_adr) X(merge_dst) X(merge_src)                 \
  X(cchild_id) X(hpass_desc) X(apass_desc) X(seg_end) X(seg_cb)                                   \
  X(cgeom_bodyid) X(con_g1) X(con_g2) X(con_cb1) X(con_cb2) X(con_ref) X(con_fn) X(con_sub) X(con_dim) X(con_xref) X(con_xslot) X(con_seg) \
  X(cbcon_adr) X(cbcon_cs) X(dof_wgrp) X(wgrp_adr) X(wgrp_cb) X(wgrp_rng) X(cb_lastdof)                                     \
  X(joint_idxs) X(body_idxs) X(endeff_idxs) X(animal_rec) X(jidx_adr) X(bidx_adr) X(eidx_adr)

// ---- float tables ----
#define BT_FLT_TABLES(X) \
  X(qpos0) X(dof_armature) X(dof_damping) X(dof_range) X(dof_solref) X(dof_solimp) X(dof_margin) X(dof_invweight0) \
  X(cgeom_pos) X(cgeom_quat) X(cgeom_size)                                                                      \
  X(con_mu) X(con_solref) X(con_solimp) X(con_includemargin) X(con_invweight)                                   \
  X(act_rec) X(wrap_rec) X(dof_rec) X(dofact_rec) X(body_rec) X(jnt_rec) X(bl_rec) X(sh_tab)                                                    \
  X(clip_position) X(clip_quaternion) X(clip_joints) X(clip_body_positions) X(clip_angular_velocity)

struct BtDev {
#define X(n) int n;
  BT_INT_SCALARS(X)
#undef X
#define X(n) float n;
  BT_FLT_SCALARS(X)
#undef X
#define X(n) const int* n;
  BT_INT_TABLES(X)
#undef X
#define X(n) const float* n;
  BT_FLT_TABLES(X)
#undef X
};

enum { BT_JNT_FREE = 0, BT_JNT_HINGE = 3 };
enum { BT_CONE_PYRAMIDAL = 0, BT_CONE_ELLIPTIC = 1 };
enum { BT_FN_PLANE_CAPSULE = 0, BT_FN_PLANE_ELLIPSOID = 1, BT_FN_PLANE_SPHERE = 2, BT_FN_CAPSULE_CAPSULE = 3 };

// per-env state rows handed through the C ABI: BtStatePtrs in include/bt_api.h
#include "bt_api.h"
typedef BtStatePtrs BtState;

// metrics order (fruitfly.py:481-494)
enum {
  BT_M_POS = 0, BT_M_QUAT, BT_M_JOINT, BT_M_ANGVEL, BT_M_BODYPOS, BT_M_ENDEFF, BT_M_QUADCTRL, BT_M_ALIVE,
  BT_M_TOO_FAR, BT_M_BAD_POSE, BT_M_BAD_QUAT, BT_M_FALL, BT_NMETRIC
};
// info_f order: summed_pos_distance, quat_distance, joint_distance, steps (float, EpisodeWrapper), truncation
enum { BT_IF_SUMMED_POS = 0, BT_IF_QUAT, BT_IF_JOINT, BT_IF_STEPS, BT_IF_TRUNC, BT_NINFOF };
// info_i order: cur_frame, steps_taken_cur_frame
enum { BT_II_CUR_FRAME = 0, BT_II_STEPS_TAKEN, BT_NINFOI };

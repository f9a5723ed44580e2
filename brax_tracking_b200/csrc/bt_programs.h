// Per-environment programs: what one lane group executes for one environment in each C-ABI entry point.
// Shared verbatim by the sm_100a kernels (bt_kernels.cu, G = 32) and the host emulation used by the CPU
// tests (tests/host_emu, G = 1).
#pragma once
#include <stdint.h>

#include "bt_impl.h"

struct BtStepArgs {
  const float* action;  // [n, nu]
  BtState state, first;
  const float* first_obs;
  const int32_t* first_info_i;
  float *obs, *reward, *done, *metrics, *info_f;
  int32_t* info_i;
  const int32_t* clip_idx;  // [n] reference clip of every environment, or NULL (single clip)
};

// Observation row -> the caller's [n, obs_size] buffer (the largest algorithmic HBM write of the step: 2.5 - 5.2 kB per environment).
// Rows of the reference's row-major layout are 4-byte aligned only (617 floats = 2468 B per rodent row), which rules out bulk /
// TMA copies (16-byte aligned addresses AND sizes); the row is instead staged in shared memory at the SAME offset modulo 16 bytes as
// its destination (BtEnv::obs_pad), so that after 0-3 head floats the body moves with 128-bit loads and coalesced 128-bit stores
// (a quarter of the store instructions; full 16-byte write transactions, which also matters when the buffer is mapped host memory).
template <int G, int DS, int CS>
BT_DEV void bt_write_obs(BtEnv<G, DS, CS>& E, float* obs_row, const float* src) {
  const int n = E.m.obs_size;
#ifdef __CUDACC__
  if ((((uintptr_t)obs_row ^ (uintptr_t)src) & 15) == 0) {
    int head = (int)((16 - ((uintptr_t)obs_row & 15)) & 15) >> 2;
    head = head < n ? head : n;
    const int nvec = (n - head) >> 2, tail = head + 4 * nvec;
    if (E.lane < head) obs_row[E.lane] = src[E.lane];
    const float4* s4 = reinterpret_cast<const float4*>(src + head);
    float4* d4 = reinterpret_cast<float4*>(obs_row + head);
    for (int i = E.lane; i < nvec; i += G) d4[i] = s4[i];
    if (E.lane < n - tail) obs_row[tail + E.lane] = src[tail + E.lane];
    return;
  }
#endif
  for (int i = E.lane; i < n; i += G) obs_row[i] = src[i];
}

// wrap(env).step  (custom_wrappers.py:54-80 o EpisodeWrapper.step o fruitfly.py:497-596)
template <int G, int DS, int CS>
// `live` = false: this lane group has no environment in this round and only takes part in the CTA barriers of substep()
BT_DEV void bt_prog_step(const BtDev& m, float* s, int lane, int env, bool live, const BtStepArgs& a) {
  typedef BtLanes<G> W;
  BtEnv<G, DS, CS> E(m, s, lane, live);
  float steps = 0.f, time = 0.f;
  int cur_in = 0, stc_in = 0;
  if (live) {
    // AutoResetWrapperTracking.step: steps <- 0 where the previous step was done (custom_wrappers.py:55-58)
    const float prev_done = a.done[env];
    steps = a.info_f[(size_t)env * BT_NINFOF + BT_IF_STEPS];
    if (prev_done > 0.f) steps = 0.f;
    cur_in = a.info_i[(size_t)env * BT_NINFOI + BT_II_CUR_FRAME];
    stc_in = a.info_i[(size_t)env * BT_NINFOI + BT_II_STEPS_TAKEN];
    time = a.state.time[env];
    if (a.clip_idx) E.clip = bt_clampi(a.clip_idx[env], 0, m.n_clips - 1);
    W::sync();  // all lanes have read done / info before lane 0 overwrites them below
    E.poison_scratch();
    E.load_state(a.state, env);
    const float* act_row = a.action + (size_t)env * m.nu;
    for (int u = lane; u < m.nu; u += G) E.ctrl()[u] = act_row[u];
    W::sync();
  }
  for (int f = 0; f < m.n_frames; f++) {
    E.step(f);
    time += m.timestep;
  }
  if (!live) return;
  typename BtEnv<G, DS, CS>::StepOut r;
  E.reward_terms(E.ctrl(), cur_in, stc_in, r);
  float* obs_row = a.obs + (size_t)env * m.obs_size;
  E.obs_pad = (int)(((uintptr_t)obs_row >> 2) & 3);
  E.build_obs(r.cur_frame);
  // EpisodeWrapper.step (action_repeat = 1)
  steps += 1.f;
  const bool over = steps >= (float)m.episode_length;
  const float done = over ? 1.f : r.done;
  const float trunc = over ? 1.f - r.done : 0.f;
  int cur = r.cur_frame, stc = r.steps_taken;
  if (done > 0.f) {
    // reset selection (custom_wrappers.py:62-80): restore the cached first state / obs / frame counters
    const size_t e = (size_t)env;
    for (int i = lane; i < m.nq; i += G) a.state.qpos[e * m.nq + i] = a.first.qpos[e * m.nq + i];
    for (int i = lane; i < m.nv; i += G) {
      a.state.qvel[e * m.nv + i] = a.first.qvel[e * m.nv + i];
      a.state.qacc_warmstart[e * m.nv + i] = a.first.qacc_warmstart[e * m.nv + i];
    }
    for (int i = lane; i < m.na; i += G) a.state.act[e * m.na + i] = a.first.act[e * m.na + i];
    if (a.state.xpos) for (int i = lane; i < 3 * m.nbody; i += G) a.state.xpos[e * 3 * m.nbody + i] = a.first.xpos[e * 3 * m.nbody + i];
    if (lane == 0) a.state.time[env] = a.first.time[env];
    bt_write_obs(E, obs_row, a.first_obs + e * m.obs_size);
    cur = a.first_info_i[e * BT_NINFOI + BT_II_CUR_FRAME];
    stc = a.first_info_i[e * BT_NINFOI + BT_II_STEPS_TAKEN];
  } else {
    E.store_state(a.state, env, time);
    bt_write_obs(E, obs_row, E.obsbuf());
  }
  if (lane == 0) {
    a.reward[env] = r.reward;
    a.done[env] = done;
    float* mt = a.metrics + (size_t)env * BT_NMETRIC;
#pragma unroll
    for (int k = 0; k < BT_NMETRIC; k++) mt[k] = r.metrics[k];
    float* nf = a.info_f + (size_t)env * BT_NINFOF;
    nf[BT_IF_SUMMED_POS] = r.summed_pos; nf[BT_IF_QUAT] = r.quat_d; nf[BT_IF_JOINT] = r.joint_d;
    nf[BT_IF_STEPS] = steps; nf[BT_IF_TRUNC] = trunc;
    int32_t* ni = a.info_i + (size_t)env * BT_NINFOI;
    ni[BT_II_CUR_FRAME] = cur; ni[BT_II_STEPS_TAKEN] = stc;
  }
  W::sync();
}

struct BtResetArgs {
  const uint32_t* keys;  // [n, 2]
  int fixed_start_frame; // < 0: training reset (fruitfly.py:449-495); >= 0: RenderRolloutWrapperTracking.reset (custom_wrappers.py:85-125)
  BtState state;
  float *obs, *reward, *done, *metrics, *info_f;
  int32_t* info_i;
  int32_t* clip_idx;  // [n] OUT: the clip drawn for every environment (randint(rng_pos, (), 0, n_clips)), or NULL (single clip)
};

// wrap(env).reset  (fruitfly.py:449-495, rodent.py:154-159; JAX threefry per SURVEY Appendix D)
template <int G, int DS, int CS>
BT_DEV void bt_prog_reset(const BtDev& m, float* s, int lane, int env, bool live, const BtResetArgs& a) {
  typedef BtLanes<G> W;
  BtEnv<G, DS, CS> E(m, s, lane, live);
  if (!live) { E.forward(); return; }
  E.poison_scratch();
  const unsigned k0 = a.keys[2 * (size_t)env], k1 = a.keys[2 * (size_t)env + 1];
  // training: rng, rng1, rng2, rng_pos = split(rng, 4);  render rollout: rng, rng1, rng2 = split(rng, 3)
  const int nsplit = a.fixed_start_frame < 0 ? 8 : 6;
  unsigned sk[8];
#pragma unroll
  for (int i = 0; i < 8; i++) sk[i] = i < nsplit ? bt_random_bits(k0, k1, i, nsplit) : 0u;
  // start_frame = randint(rng, (), 0, range): k1_, k2_ = split(rng); bits of each; span arithmetic in uint32
  const unsigned r0 = sk[0], r1 = sk[1];
  unsigned ss[4];
#pragma unroll
  for (int i = 0; i < 4; i++) ss[i] = bt_random_bits(r0, r1, i, 4);
  const unsigned hi_bits = bt_random_bits(ss[0], ss[1], 0, 1), lo_bits = bt_random_bits(ss[2], ss[3], 0, 1);
  unsigned span = (unsigned)m.start_frame_range;
  if (span == 0) span = 1;
  unsigned mult = 65536u % span;
  mult = (mult * mult) % span;
  const unsigned off = (hi_bits % span) * mult + (lo_bits % span);
  const int start = a.fixed_start_frame < 0 ? (int)(off % span) : a.fixed_start_frame;
  // RodentMultiClip (envs/rodent.py:377, an empty class in the reference): the clip of an environment
  // is drawn from the fourth key of the reset's split(rng, 4) -- `rng_pos`, which the reference draws and never uses
  // (envs/fruitfly.py:451) -- as randint(rng_pos, (), 0, n_clips); it stays with the environment through auto-resets, which
  // restore the cached first state (custom_wrappers.py:62-80)
  if (a.clip_idx && m.n_clips > 1 && a.fixed_start_frame < 0) {
    unsigned cs4[4];
#pragma unroll
    for (int i = 0; i < 4; i++) cs4[i] = bt_random_bits(sk[6], sk[7], i, 4);
    const unsigned chi = bt_random_bits(cs4[0], cs4[1], 0, 1), clo = bt_random_bits(cs4[2], cs4[3], 0, 1);
    const unsigned cspan = (unsigned)m.n_clips;
    unsigned cmult = 65536u % cspan;
    cmult = (cmult * cmult) % cspan;
    E.clip = (int)(((chi % cspan) * cmult + (clo % cspan)) % cspan);
  } else if (a.clip_idx && m.n_clips > 1) {
    // render / evaluation rollout (no rng_pos in its split(rng, 3)): the caller chooses the clip
    E.clip = bt_clampi(a.clip_idx[env], 0, m.n_clips - 1);
  }
  const int crow = E.clip * m.clip_len;
  const float lo = -m.reset_noise_scale, hi = m.reset_noise_scale;
  for (int i = lane; i < m.nq; i += G) {
    float q0 = BT_LDG(m.qpos0 + i);
    if (m.seed_root_from_clip && a.fixed_start_frame < 0) {  // the render-rollout reset starts from qpos0 (custom_wrappers.py:99-103)
      const int fs = start > m.clip_len - 1 ? m.clip_len - 1 : start;  // JAX gather clamps (a clip shorter than the start-frame range)
      for (int an = 0; an < m.n_animals; an++) {   // every animal's root x, y and orientation from its own copy of the clip
        const int rel = i - BT_LDG(m.animal_rec + 8 * an);
        if (rel >= 0 && rel < 2) q0 = BT_LDG(m.clip_position + 3 * ((crow + fs) * m.n_animals + an) + rel);
        else if (rel >= 3 && rel < 7) q0 = BT_LDG(m.clip_quaternion + 4 * ((crow + fs) * m.n_animals + an) + rel - 3);
      }
    }
    E.qpos()[i] = q0 + bt_bits_to_uniform(bt_random_bits(sk[2], sk[3], i, m.nq), lo, hi);
  }
  for (int i = lane; i < m.nv; i += G) {
    E.qvel()[i] = bt_bits_to_uniform(bt_random_bits(sk[4], sk[5], i, m.nv), lo, hi);
    E.warm()[i] = 0.f;
  }
  for (int i = lane; i < m.na; i += G) E.act()[i] = 0.f;
  for (int i = lane; i < m.nu; i += G) E.ctrl()[i] = 0.f;
  W::sync();
  E.forward();  // pipeline_init = mjx.forward (leaves qacc_warmstart = qacc)
  E.obs_pad = (int)(((uintptr_t)(a.obs + (size_t)env * m.obs_size) >> 2) & 3);
  E.build_obs(start);
  E.store_state(a.state, env, 0.f);
  bt_write_obs(E, a.obs + (size_t)env * m.obs_size, E.obsbuf());
  if (lane == 0) {
    a.reward[env] = 0.f;
    a.done[env] = 0.f;
    for (int k = 0; k < BT_NMETRIC; k++) a.metrics[(size_t)env * BT_NMETRIC + k] = 0.f;
    for (int k = 0; k < BT_NINFOF; k++) a.info_f[(size_t)env * BT_NINFOF + k] = 0.f;
    a.info_i[(size_t)env * BT_NINFOI + BT_II_CUR_FRAME] = start;
    a.info_i[(size_t)env * BT_NINFOI + BT_II_STEPS_TAKEN] = 0;
    if (a.clip_idx) a.clip_idx[env] = E.clip;
  }
  W::sync();
}

// PipelineEnv.pipeline_step (n_substeps > 0) or pipeline_init (n_substeps == 0: one mjx.forward)
template <int G, int DS, int CS>
BT_DEV void bt_prog_physics(const BtDev& m, float* s, int lane, int env, bool live, const float* ctrl, const BtState& st,
                            int n_substeps) {
  typedef BtLanes<G> W;
  BtEnv<G, DS, CS> E(m, s, lane, live);
  float time = 0.f;
  if (live) {
    time = st.time[env];
    W::sync();
    E.poison_scratch();
    E.load_state(st, env);
    for (int u = lane; u < m.nu; u += G) E.ctrl()[u] = ctrl ? ctrl[(size_t)env * m.nu + u] : 0.f;
    W::sync();
  }
  if (n_substeps == 0) E.forward();
  for (int f = 0; f < n_substeps; f++) {
    E.step();
    time += m.timestep;
  }
  if (live) E.store_state(st, env, time);
  W::sync();
}

struct BtRewardArgs {
  const float* action;
  BtState state;
  int32_t* info_i;
  float *obs, *reward, *done, *metrics, *info_f;
  const int32_t* clip_idx;
};

// env.step after pipeline_step (fruitfly.py:502-596), no wrappers
template <int G, int DS, int CS>
BT_DEV void bt_prog_reward(const BtDev& m, float* s, int lane, int env, bool live, const BtRewardArgs& a) {
  typedef BtLanes<G> W;
  if (!live) return;
  BtEnv<G, DS, CS> E(m, s, lane);
  const int cur_in = a.info_i[(size_t)env * BT_NINFOI + BT_II_CUR_FRAME];
  const int stc_in = a.info_i[(size_t)env * BT_NINFOI + BT_II_STEPS_TAKEN];
  if (a.clip_idx) E.clip = bt_clampi(a.clip_idx[env], 0, m.n_clips - 1);
  W::sync();
  E.poison_scratch();
  E.load_state(a.state, env);
  for (int i = lane; i < 3 * m.nbody; i += G) E.xpos()[i] = a.state.xpos[(size_t)env * 3 * m.nbody + i];
  for (int u = lane; u < m.nu; u += G) E.ctrl()[u] = a.action[(size_t)env * m.nu + u];
  W::sync();
  typename BtEnv<G, DS, CS>::StepOut r;
  E.reward_terms(E.ctrl(), cur_in, stc_in, r);
  E.obs_pad = (int)(((uintptr_t)(a.obs + (size_t)env * m.obs_size) >> 2) & 3);
  E.build_obs(r.cur_frame);
  bt_write_obs(E, a.obs + (size_t)env * m.obs_size, E.obsbuf());
  if (lane == 0) {
    a.reward[env] = r.reward;
    a.done[env] = r.done;
    for (int k = 0; k < BT_NMETRIC; k++) a.metrics[(size_t)env * BT_NMETRIC + k] = r.metrics[k];
    float* nf = a.info_f + (size_t)env * BT_NINFOF;
    nf[BT_IF_SUMMED_POS] = r.summed_pos; nf[BT_IF_QUAT] = r.quat_d; nf[BT_IF_JOINT] = r.joint_d;
    a.info_i[(size_t)env * BT_NINFOI + BT_II_CUR_FRAME] = r.cur_frame;
    a.info_i[(size_t)env * BT_NINFOI + BT_II_STEPS_TAKEN] = r.steps_taken;
  }
  W::sync();
}

// mjx.forward up to a stop point, then dump the scratch block (parity tests of intermediates)
template <int G, int DS, int CS>
BT_DEV void bt_prog_debug(const BtDev& m, float* s, int lane, int env, bool live, const float* ctrl, const BtState& st, int stop,
                          float* scratch, float* cdist, int32_t* niter) {
  typedef BtLanes<G> W;
  BtEnv<G, DS, CS> E(m, s, lane, live);
  if (!live) { E.forward(stop); return; }
  for (int i = lane; i < m.smem_floats; i += G) s[i] = 0.f;
  W::sync();
  E.load_state(st, env);
  for (int u = lane; u < m.nu; u += G) E.ctrl()[u] = ctrl ? ctrl[(size_t)env * m.nu + u] : 0.f;
#pragma unroll
  for (int sl = 0; sl < CS; sl++) E.cdist[sl] = 0.f;
  W::sync();
  E.forward(stop);
  W::sync();
  for (int i = lane; i < m.smem_floats; i += G) scratch[(size_t)env * m.smem_floats + i] = s[i];
#pragma unroll
  for (int sl = 0; sl < CS; sl++) {
    const int c = lane + sl * G;
    if (c < m.ncon) cdist[(size_t)env * m.ncon + c] = E.cdist[sl];
  }
  if (lane == 0) niter[env] = E.niter;
  W::sync();
}

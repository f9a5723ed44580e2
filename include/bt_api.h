/*
 * C ABI of the B200-native tracking step (libbt_b200.so).  Plain pointers and sizes only: every array
 * argument is a DEVICE pointer to a row-major [n_envs, dim] buffer owned by the caller; kernels are enqueued
 * on the caller's stream (a cudaStream_t passed as void*), never synchronised implicitly.  All functions
 * return 0 on success or a negative BT_E* code and never throw; bt_last_error() describes the last failure
 * of the calling thread.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference checkout):
 *
 *   bt_model_create   <- the env constructor: mjx.put_model + clip / index tables
 *                        (envs/fruitfly.py:346-447, envs/rodent.py:19-136)
 *   bt_reset          <- wrap(env).reset: Fruitfly_Tethered_Free.reset (envs/fruitfly.py:449-495; rodent
 *                        root seeding envs/rodent.py:154-159) under EpisodeWrapper / VmapWrapper /
 *                        AutoResetWrapperTracking.reset (custom_brax/custom_wrappers.py:46-52)
 *   bt_step           <- wrap(env).step: AutoResetWrapperTracking.step (custom_brax/custom_wrappers.py:54-80)
 *                        o EpisodeWrapper.step o Fruitfly_Tethered_Free.step (envs/fruitfly.py:497-596),
 *                        including PipelineEnv.pipeline_step = mjx.step x n_frames (envs/fruitfly.py:500)
 *   bt_physics_step   <- PipelineEnv.pipeline_step alone (envs/fruitfly.py:500; brax.mjx.pipeline.step)
 *   bt_pipeline_init  <- PipelineEnv.pipeline_init = mjx.forward (envs/fruitfly.py:477)
 *   bt_reward_obs     <- the part of env.step after pipeline_step (envs/fruitfly.py:502-596, _get_obs :598-646)
 *   bt_ppo_tanh_normal_fwd / _bwd <- the policy terms of brax compute_ppo_loss (custom_brax/custom_ppo.py:250-284)
 *   bt_ppo_flat_adam  <- optax.adam + the mean of lax.pmean(grads) (custom_brax/custom_ppo.py:233,246-257)
 *   bt_forward_debug  <- no reference counterpart: dumps on-chip intermediates for the parity tests
 */
#ifndef BT_API_H_
#define BT_API_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct BtModel BtModel;

enum {
  BT_OK = 0,
  BT_E_ARG = -1,      /* bad argument / missing or mis-sized table */
  BT_E_CUDA = -2,     /* CUDA runtime error (see bt_last_error) */
  BT_E_UNSUPPORTED = -3 /* model exceeds the compiled kernel variants */
};

/* per-env state rows (pipeline_state fields read outside the physics: SURVEY.md section 8b) */
typedef struct BtStatePtrs {
  float* qpos;           /* [n, nq] */
  float* qvel;           /* [n, nv] */
  float* act;            /* [n, na]  (may be NULL when na == 0) */
  float* qacc_warmstart; /* [n, nv] */
  float* time;           /* [n] */
  float* xpos;           /* [n, nbody*3] */
} BtStatePtrs;

#define BT_NUM_METRICS 12 /* fruitfly.py:481-494 order */
#define BT_NUM_INFO_F 5   /* summed_pos_distance, quat_distance, joint_distance, steps, truncation */
#define BT_NUM_INFO_I 2   /* cur_frame, steps_taken_cur_frame */

/* tables: n_tables named host arrays (int32 or float32) as produced by model.py::pack */
int bt_model_create(int n_tables, const char* const* names, const void* const* data, const int64_t* counts,
                    const int* is_float, int device, BtModel** out);
void bt_model_destroy(BtModel* m);
/* dims[0..8] = nq, nv, nu, na, nbody, obs_size, smem_floats, ncon, n_clips */
int bt_model_dims(const BtModel* m, int* dims);
/* launch geometry chosen for this model: out[0] = warps (= envs) per CTA, out[1] = max CTAs, out[2] = dynamic smem bytes */
int bt_model_launch(const BtModel* m, int* out);

/* fixed_start_frame < 0: training reset (random start frame in [0, 44), split(rng, 4));
   fixed_start_frame >= 0: RenderRolloutWrapperTracking.reset (custom_brax/custom_wrappers.py:85-125: that frame, split(rng, 3)) */
/* clip_idx [n] (may be NULL: single clip): the reference clip of every environment when the model holds several stacked clips
   (RodentMultiClip, envs/rodent.py:377; preprocessing/preprocess.py:254-258).  bt_reset WRITES it on a training reset
   (randint(rng_pos, (), 0, n_clips), rng_pos = the fourth key of the reset's split(rng, 4)) and READS it on a render reset;
   bt_step / bt_reward_obs read it.  It stays with the environment through auto-resets, like the cached first state. */
int bt_reset(BtModel* m, int n_envs, const uint32_t* keys /*[n,2]*/, int fixed_start_frame, BtStatePtrs state,
             float* obs /*[n,O]*/, float* reward, float* done, float* metrics /*[n,12]*/, float* info_f /*[n,5]*/,
             int32_t* info_i /*[n,2]*/, int32_t* clip_idx /*[n] or NULL*/, void* stream);

int bt_step(BtModel* m, int n_envs, const float* action /*[n,nu]*/, BtStatePtrs state /*in-out*/,
            BtStatePtrs first /*auto-reset source*/, const float* first_obs, const int32_t* first_info_i,
            float* obs, float* reward, float* done /*in: previous, out: new*/, float* metrics, float* info_f,
            int32_t* info_i, const int32_t* clip_idx /*[n] or NULL*/, void* stream);

int bt_physics_step(BtModel* m, int n_envs, const float* ctrl /*[n,nu]*/, BtStatePtrs state, int n_substeps,
                    void* stream);
int bt_pipeline_init(BtModel* m, int n_envs, BtStatePtrs state, void* stream);
int bt_reward_obs(BtModel* m, int n_envs, const float* action, BtStatePtrs state /*read only*/,
                  int32_t* info_i /*in-out*/, float* obs, float* reward, float* done, float* metrics,
                  float* info_f, const int32_t* clip_idx /*[n] or NULL*/, void* stream);
/* runs mjx.forward up to `stop` (0 = all of it) and copies every env's scratch block: scratch [n, smem_floats],
   cdist [n, ncon], niter [n] */
int bt_forward_debug(BtModel* m, int n_envs, const float* ctrl, BtStatePtrs state, int stop, float* scratch,
                     float* cdist, int32_t* niter, void* stream);

/* Learner-side helper of the PPO loop that drives bt_step (custom_brax/custom_ppo.py:250-284 -> brax compute_ppo_loss with
   NormalTanhDistribution, min_std 0.001): per (b, t) row of logits [B, T, 2A] (contiguous; loc | pre-softplus scale), the
   log-probability of the stored raw action and the sampled-entropy term, summed over the A action dims, in one pass; and the
   matching gradient w.r.t. the logits.  raw / noise [.., A] and the per-row outputs / output gradients are addressed as
   b * sb + t * st (element strides), innermost dim contiguous. */
int bt_ppo_tanh_normal_fwd(int B, int T, int A, const float* logits, const float* raw, int64_t raw_sb, int64_t raw_st,
                           const float* noise, int64_t noise_sb, int64_t noise_st, float* lp, float* ent, int64_t out_sb,
                           int64_t out_st, void* stream);
int bt_ppo_tanh_normal_bwd(int B, int T, int A, const float* logits, const float* raw, int64_t raw_sb, int64_t raw_st,
                           const float* noise, int64_t noise_sb, int64_t noise_st, const float* glp, const float* gent,
                           int64_t out_sb, int64_t out_st, float* glogits, void* stream);

/* optax.adam on a flat parameter buffer fused with the 1/world scale of lax.pmean (custom_brax/custom_ppo.py:233,246-257): p, g,
   m, v are [n] device floats, `step` a device float with the number of updates already applied (the caller advances it). */
int bt_ppo_flat_adam(int64_t n, float* p, const float* g, float* m, float* v, const float* step, float lr, float b1, float b2,
                     float eps, float gscale, void* stream);

const char* bt_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t bt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif

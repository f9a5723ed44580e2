// C ABI (include/bt_api.h) of the fused physics + tracking-reward step: model upload, variant selection and
// kernel launches.  The kernels themselves are in bt_tu.inc (one translation unit per variant and kernel).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "bt_api.h"
#include "bt_bind.h"
#include "bt_ops.h"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

// (dof slots per lane, contact slots per lane): rodent nv 73 / 30 contacts -> (3,1); fly nv 42|36 / 27 -> (2,1);
// two-rodent stress config nv 146 / 114 contacts -> (5,4).  Keep in sync with VARIANTS in build.py.
#define BT_VARIANTS(X) X(2, 1) X(3, 1) X(5, 4)

#define X(d, c)                                                                                                     \
  cudaError_t bt_prepare_step_##d##_##c(int); cudaError_t bt_prepare_reset_##d##_##c(int);                          \
  cudaError_t bt_prepare_physics_##d##_##c(int); cudaError_t bt_prepare_reward_##d##_##c(int);                      \
  cudaError_t bt_prepare_debug_##d##_##c(int);                                                                      \
  void bt_launch_step_##d##_##c(BtLaunchCfg, const BtDev&, int, int, const BtStepArgs&);                            \
  void bt_launch_reset_##d##_##c(BtLaunchCfg, const BtDev&, int, int, const BtResetArgs&);                          \
  void bt_launch_physics_##d##_##c(BtLaunchCfg, const BtDev&, int, int, const float*, const BtState&, int);         \
  void bt_launch_reward_##d##_##c(BtLaunchCfg, const BtDev&, int, int, const BtRewardArgs&);                        \
  void bt_launch_debug_##d##_##c(BtLaunchCfg, const BtDev&, int, int, const float*, const BtState&, int, float*, float*, int32_t*);
BT_VARIANTS(X)
#undef X

static const BtVariantOps kVariants[] = {
#define X(d, c)                                                                                                     \
  {d, c, BT_VARIANT_MAX_WARPS(d, c), bt_prepare_step_##d##_##c, bt_prepare_reset_##d##_##c, bt_prepare_physics_##d##_##c, bt_prepare_reward_##d##_##c, \
   bt_prepare_debug_##d##_##c, bt_launch_step_##d##_##c, bt_launch_reset_##d##_##c, bt_launch_physics_##d##_##c,      \
   bt_launch_reward_##d##_##c, bt_launch_debug_##d##_##c},
    BT_VARIANTS(X)
#undef X
};

struct BtModel {
  BtDev dev;        // scalars + device table pointers
  void* blob;       // one allocation holding every table
  int device;
  const BtVariantOps* ops;
  int warps;        // environments per CTA
  int max_ctas;     // persistent grid size cap
  int smem_bytes;   // dynamic shared memory per CTA
};

#define BT_CUDA(call)                                                                           \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      snprintf(g_err, sizeof(g_err), "%s failed: %s", #call, cudaGetErrorString(e_));           \
      return BT_E_CUDA;                                                                         \
    }                                                                                           \
  } while (0)

// Every entry point that touches the GPU runs with the model's device current and restores the caller's device on exit
// (the stream handed in must belong to that device; cudaErrorInvalidResourceHandle otherwise, reported by BT_LAUNCHED).
struct BtDeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit BtDeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) { err = cudaSetDevice(device); switched = err == cudaSuccess; }
  }
  ~BtDeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define BT_ON_DEVICE(m)                                                                                          \
  BtDeviceGuard guard_((m)->device);                                                                             \
  if (guard_.err != cudaSuccess) {                                                                               \
    snprintf(g_err, sizeof(g_err), "cudaSetDevice(%d) failed: %s", (m)->device, cudaGetErrorString(guard_.err)); \
    return BT_E_CUDA;                                                                                            \
  }

static inline BtLaunchCfg cfg_for(const BtModel* m, int n_envs, void* stream) {
  // Persistent CTAs of `warps` environments each.  A batch that does not fill every SM at the full width (a strong-scaling shard,
  // an evaluation batch) is spread over ALL the SMs with narrower CTAs instead of leaving SMs idle: warps = ceil(n / SMs).
  int warps = m->warps;
  if ((int64_t)n_envs < (int64_t)m->max_ctas * m->warps) {
    warps = (n_envs + m->max_ctas - 1) / m->max_ctas;
    if (warps < 1) warps = 1;
  }
  int ctas = (n_envs + warps - 1) / warps;
  if (ctas > m->max_ctas) ctas = m->max_ctas;
  if (ctas < 1) ctas = 1;
  BtLaunchCfg c = {ctas, warps * 32, (int)((size_t)m->dev.sh_stage_floats * 4 + (size_t)m->dev.smem_floats * 4 * warps), (cudaStream_t)stream, warps};
  return c;
}
#define BT_LAUNCHED()                    \
  do {                                   \
    BT_CUDA(cudaGetLastError());         \
    g_launches++;                        \
  } while (0)

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

const char* bt_last_error(void) { return g_err; }
int64_t bt_launch_count(void) { return g_launches.load(); }

int bt_model_create(int n, const char* const* names, const void* const* data, const int64_t* counts, const int* is_float,
                    int device, BtModel** out) {
  if (!names || !data || !counts || !is_float || !out || n <= 0) { snprintf(g_err, sizeof(g_err), "null argument"); return BT_E_ARG; }
  BtDeviceGuard guard(device);
  if (guard.err != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(guard.err)); return BT_E_CUDA; }
  // one blob, every table 256-byte aligned
  std::vector<size_t> off(n);
  size_t total = 0;
  for (int i = 0; i < n; i++) {
    if (counts[i] <= 0 || !data[i]) { snprintf(g_err, sizeof(g_err), "table '%s' is empty", names[i]); return BT_E_ARG; }
    off[i] = total;
    total += ((size_t)counts[i] * 4 + 255) & ~(size_t)255;
  }
  std::vector<char> host(total, 0);
  for (int i = 0; i < n; i++) memcpy(host.data() + off[i], data[i], (size_t)counts[i] * 4);
  BtModel* m = new BtModel();
  m->blob = nullptr;
  m->device = device;
  // every error path below goes through fail(): frees the blob and the handle (the guard restores the caller's device)
  auto fail = [&](int code) { if (m->blob) cudaFree(m->blob); delete m; return code; };
  cudaError_t e = cudaMalloc(&m->blob, total);
  if (e != cudaSuccess) { m->blob = nullptr; snprintf(g_err, sizeof(g_err), "cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e)); return fail(BT_E_CUDA); }
  e = cudaMemcpy(m->blob, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "upload failed: %s", cudaGetErrorString(e)); return fail(BT_E_CUDA); }
  std::vector<const void*> bound(n);
  for (int i = 0; i < n; i++) bound[i] = (const char*)m->blob + off[i];
  if (bt_bind(&m->dev, n, names, data, bound.data(), counts, is_float, g_err, sizeof(g_err))) return fail(BT_E_ARG);
  const int need_ds = (m->dev.nv + 31) / 32, need_cs = m->dev.ncon > 0 ? (m->dev.ncon + 31) / 32 : 1;
  m->ops = nullptr;
  for (const BtVariantOps& v : kVariants)
    if (v.ds >= need_ds && v.cs >= need_cs) { m->ops = &v; break; }
  if (!m->ops) {
    snprintf(g_err, sizeof(g_err), "model (nv=%d, ncon=%d) exceeds the compiled kernel variants", m->dev.nv, m->dev.ncon);
    return fail(BT_E_UNSUPPORTED);
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e)); return fail(BT_E_CUDA); }
  const size_t per_env = (size_t)m->dev.smem_floats * 4;
  const size_t staged = (size_t)m->dev.sh_stage_floats * 4;   // CTA-shared constant records in front of the environments' slices
  if (staged > prop.sharedMemPerBlockOptin || m->dev.sh_stage_floats < 0 || (m->dev.sh_stage_floats & 3)) {
    snprintf(g_err, sizeof(g_err), "sh_stage_floats = %d is not a multiple of 4 within shared memory", m->dev.sh_stage_floats);
    return fail(BT_E_ARG);
  }
  int warps = (int)((prop.sharedMemPerBlockOptin - staged) / per_env);
  if (warps > m->ops->max_warps) warps = m->ops->max_warps;
  if (warps < 1) {
    snprintf(g_err, sizeof(g_err), "per-environment scratch (%zu B) exceeds shared memory", per_env);
    return fail(BT_E_UNSUPPORTED);
  }
  if (const char* w = getenv("BT_WARPS")) { int v = atoi(w); if (v >= 1 && v <= warps) warps = v; }
  m->warps = warps;
  m->max_ctas = prop.multiProcessorCount;
  m->smem_bytes = (int)(staged + per_env * warps);
  {
    const BtVariantOps* o = m->ops;
    cudaError_t pe = o->prepare_step(m->smem_bytes);
    if (pe == cudaSuccess) pe = o->prepare_reset(m->smem_bytes);
    if (pe == cudaSuccess) pe = o->prepare_physics(m->smem_bytes);
    if (pe == cudaSuccess) pe = o->prepare_reward(m->smem_bytes);
    if (pe == cudaSuccess) pe = o->prepare_debug(m->smem_bytes);
    if (pe != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(%d B dynamic smem) failed: %s", m->smem_bytes, cudaGetErrorString(pe));
      return fail(BT_E_CUDA);
    }
  }
  *out = m;
  return BT_OK;
}

void bt_model_destroy(BtModel* m) {
  if (!m) return;
  BtDeviceGuard guard(m->device);
  cudaFree(m->blob);
  delete m;
}

int bt_model_dims(const BtModel* m, int* dims) {
  if (!m || !dims) { snprintf(g_err, sizeof(g_err), "null argument"); return BT_E_ARG; }
  dims[0] = m->dev.nq; dims[1] = m->dev.nv; dims[2] = m->dev.nu; dims[3] = m->dev.na; dims[4] = m->dev.nbody;
  dims[5] = m->dev.obs_size; dims[6] = m->dev.smem_floats; dims[7] = m->dev.ncon; dims[8] = m->dev.n_clips;
  return BT_OK;
}

int bt_model_launch(const BtModel* m, int* out) {
  if (!m || !out) { snprintf(g_err, sizeof(g_err), "null argument"); return BT_E_ARG; }
  out[0] = m->warps; out[1] = m->max_ctas; out[2] = m->smem_bytes;
  return BT_OK;
}

static int check_state(const BtModel* m, const BtStatePtrs& s, bool need_xpos) {
  if (!s.qpos || !s.qvel || !s.qacc_warmstart || !s.time || (m->dev.na > 0 && !s.act) || (need_xpos && !s.xpos)) {
    snprintf(g_err, sizeof(g_err), "null state pointer");
    return BT_E_ARG;
  }
  return 0;
}

int bt_reset(BtModel* m, int n_envs, const uint32_t* keys, int fixed_start_frame, BtStatePtrs state, float* obs, float* reward, float* done,
             float* metrics, float* info_f, int32_t* info_i, int32_t* clip_idx, void* stream) {
  if (!m || n_envs < 0 || !keys || !obs || !reward || !done || !metrics || !info_f || !info_i) { snprintf(g_err, sizeof(g_err), "bad argument"); return BT_E_ARG; }
  if (check_state(m, state, true)) return BT_E_ARG;
  if (n_envs == 0) return BT_OK;
  BT_ON_DEVICE(m);
  if (fixed_start_frame >= m->dev.clip_len) { snprintf(g_err, sizeof(g_err), "fixed_start_frame beyond the clip"); return BT_E_ARG; }
  BtResetArgs a = {keys, fixed_start_frame, state, obs, reward, done, metrics, info_f, info_i, clip_idx};
  { BtLaunchCfg c = cfg_for(m, n_envs, stream); m->ops->reset(c, m->dev, n_envs, c.warps, a); }
  BT_LAUNCHED();
  return BT_OK;
}

int bt_step(BtModel* m, int n_envs, const float* action, BtStatePtrs state, BtStatePtrs first, const float* first_obs,
            const int32_t* first_info_i, float* obs, float* reward, float* done, float* metrics, float* info_f, int32_t* info_i,
            const int32_t* clip_idx, void* stream) {
  if (!m || n_envs < 0 || !action || !first_obs || !first_info_i || !obs || !reward || !done || !metrics || !info_f || !info_i) {
    snprintf(g_err, sizeof(g_err), "bad argument");
    return BT_E_ARG;
  }
  if (check_state(m, state, true) || check_state(m, first, true)) return BT_E_ARG;
  if (n_envs == 0) return BT_OK;
  BT_ON_DEVICE(m);
  BtStepArgs a = {action, state, first, first_obs, first_info_i, obs, reward, done, metrics, info_f, info_i, clip_idx};
  { BtLaunchCfg c = cfg_for(m, n_envs, stream); m->ops->step(c, m->dev, n_envs, c.warps, a); }
  BT_LAUNCHED();
  return BT_OK;
}

int bt_physics_step(BtModel* m, int n_envs, const float* ctrl, BtStatePtrs state, int n_substeps, void* stream) {
  if (!m || n_envs < 0 || n_substeps < 1 || (m->dev.nu > 0 && !ctrl)) { snprintf(g_err, sizeof(g_err), "bad argument"); return BT_E_ARG; }
  if (check_state(m, state, false)) return BT_E_ARG;
  if (n_envs == 0) return BT_OK;
  BT_ON_DEVICE(m);
  { BtLaunchCfg c = cfg_for(m, n_envs, stream); m->ops->physics(c, m->dev, n_envs, c.warps, ctrl, state, n_substeps); }
  BT_LAUNCHED();
  return BT_OK;
}

int bt_pipeline_init(BtModel* m, int n_envs, BtStatePtrs state, void* stream) {
  if (!m || n_envs < 0) { snprintf(g_err, sizeof(g_err), "bad argument"); return BT_E_ARG; }
  if (check_state(m, state, false)) return BT_E_ARG;
  if (n_envs == 0) return BT_OK;
  BT_ON_DEVICE(m);
  { BtLaunchCfg c = cfg_for(m, n_envs, stream); m->ops->physics(c, m->dev, n_envs, c.warps, nullptr, state, 0); }
  BT_LAUNCHED();
  return BT_OK;
}

int bt_reward_obs(BtModel* m, int n_envs, const float* action, BtStatePtrs state, int32_t* info_i, float* obs, float* reward,
                  float* done, float* metrics, float* info_f, const int32_t* clip_idx, void* stream) {
  if (!m || n_envs < 0 || !action || !info_i || !obs || !reward || !done || !metrics || !info_f) { snprintf(g_err, sizeof(g_err), "bad argument"); return BT_E_ARG; }
  if (check_state(m, state, true)) return BT_E_ARG;
  if (n_envs == 0) return BT_OK;
  BT_ON_DEVICE(m);
  BtRewardArgs a = {action, state, info_i, obs, reward, done, metrics, info_f, clip_idx};
  { BtLaunchCfg c = cfg_for(m, n_envs, stream); m->ops->reward(c, m->dev, n_envs, c.warps, a); }
  BT_LAUNCHED();
  return BT_OK;
}

int bt_forward_debug(BtModel* m, int n_envs, const float* ctrl, BtStatePtrs state, int stop, float* scratch, float* cdist,
                     int32_t* niter, void* stream) {
  if (!m || n_envs < 0 || !scratch || !cdist || !niter) { snprintf(g_err, sizeof(g_err), "bad argument"); return BT_E_ARG; }
  if (check_state(m, state, false)) return BT_E_ARG;
  if (n_envs == 0) return BT_OK;
  BT_ON_DEVICE(m);
  { BtLaunchCfg c = cfg_for(m, n_envs, stream); m->ops->debug(c, m->dev, n_envs, c.warps, ctrl, state, stop, scratch, cdist, niter); }
  BT_LAUNCHED();
  return BT_OK;
}

}  // extern "C"

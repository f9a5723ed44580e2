// HOST EMULATION -- TEST INFRASTRUCTURE ONLY (never loaded by the product package).
//
// Compiles the exact per-environment programs of the sm_100a kernels (brax_tracking_b200/csrc/bt_programs.h)
// with one lane per environment (G = 1), so the table logic, the tree-sparse factorisation, the matrix-free
// Jacobian and the env layer can be checked against the oracle on the CPU-only box.  It exercises none of the
// warp-level synchronisation; that is what the `-m gpu` tests and compute-sanitizer are for.
#include <stdlib.h>

#include <vector>

#include "bt_bind.h"
#include "bt_programs.h"

#define EMU_DS 192
#define EMU_CS 160

struct EmuModel {
  BtDev dev;
  std::vector<std::vector<char>> keep;
};
static char g_err[256];

extern "C" {
const char* emu_last_error(void) { return g_err; }

// csrc/bt_math.h::bt_random_bits, element by element (tests/test_prng.py)
uint32_t emu_random_bits(uint32_t k0, uint32_t k1, int idx, int n) { return bt_random_bits(k0, k1, idx, n); }

int emu_model_create(int n, const char* const* names, const void* const* data, const int64_t* counts, const int* is_float,
                     EmuModel** out) {
  EmuModel* m = new EmuModel();
  std::vector<const void*> bound(n);
  for (int i = 0; i < n; i++) {
    m->keep.emplace_back((const char*)data[i], (const char*)data[i] + counts[i] * 4);
    bound[i] = m->keep.back().data();
  }
  if (bt_bind(&m->dev, n, names, data, bound.data(), counts, is_float, g_err, sizeof(g_err))) { delete m; return -1; }
  if (m->dev.nv > EMU_DS || m->dev.ncon > EMU_CS) { snprintf(g_err, sizeof(g_err), "model too large for the emulation"); delete m; return -3; }
  *out = m;
  return 0;
}
void emu_model_destroy(EmuModel* m) { delete m; }

int emu_reset(EmuModel* m, int n, const uint32_t* keys, int fixed_start_frame, BtStatePtrs st, float* obs, float* reward, float* done,
              float* metrics, float* info_f, int32_t* info_i, int32_t* clip_idx) {
  std::vector<float> s(m->dev.smem_floats);
  BtResetArgs a = {keys, fixed_start_frame, st, obs, reward, done, metrics, info_f, info_i, clip_idx};
  for (int e = 0; e < n; e++) bt_prog_reset<1, EMU_DS, EMU_CS>(m->dev, s.data(), 0, e, true, a);
  return 0;
}
int emu_step(EmuModel* m, int n, const float* action, BtStatePtrs st, BtStatePtrs first, const float* first_obs,
             const int32_t* first_info_i, float* obs, float* reward, float* done, float* metrics, float* info_f, int32_t* info_i,
             const int32_t* clip_idx) {
  std::vector<float> s(m->dev.smem_floats);
  BtStepArgs a = {action, st, first, first_obs, first_info_i, obs, reward, done, metrics, info_f, info_i, clip_idx};
  for (int e = 0; e < n; e++) bt_prog_step<1, EMU_DS, EMU_CS>(m->dev, s.data(), 0, e, true, a);
  return 0;
}
int emu_physics_step(EmuModel* m, int n, const float* ctrl, BtStatePtrs st, int n_substeps) {
  std::vector<float> s(m->dev.smem_floats);
  for (int e = 0; e < n; e++) bt_prog_physics<1, EMU_DS, EMU_CS>(m->dev, s.data(), 0, e, true, ctrl, st, n_substeps);
  return 0;
}
int emu_reward_obs(EmuModel* m, int n, const float* action, BtStatePtrs st, int32_t* info_i, float* obs, float* reward,
                   float* done, float* metrics, float* info_f, const int32_t* clip_idx) {
  std::vector<float> s(m->dev.smem_floats);
  BtRewardArgs a = {action, st, info_i, obs, reward, done, metrics, info_f, clip_idx};
  for (int e = 0; e < n; e++) bt_prog_reward<1, EMU_DS, EMU_CS>(m->dev, s.data(), 0, e, true, a);
  return 0;
}
int emu_forward_debug(EmuModel* m, int n, const float* ctrl, BtStatePtrs st, int stop, float* scratch, float* cdist, int32_t* niter) {
  std::vector<float> s(m->dev.smem_floats);
  for (int e = 0; e < n; e++) bt_prog_debug<1, EMU_DS, EMU_CS>(m->dev, s.data(), 0, e, true, ctrl, st, stop, scratch, cdist, niter);
  return 0;
}
}

// Binds named host tables (model.py::pack) to the fields of BtDev.  Shared by the CUDA library (tables are
// uploaded first; `bound[i]` is then the device address) and by the host emulation (bound[i] == host[i]).
#pragma once
#include <stdio.h>
#include <string.h>

#include "bt_model.h"

static inline int bt_find_table(int n, const char* const* names, const char* want) {
  for (int i = 0; i < n; i++)
    if (strcmp(names[i], want) == 0) return i;
  return -1;
}

// returns 0 on success; on failure writes a message into err
static inline int bt_bind(BtDev* d, int n, const char* const* names, const void* const* host, const void* const* bound,
                          const int64_t* counts, const int* is_float, char* err, size_t errlen) {
  memset(d, 0, sizeof(*d));
  int idx;
#define BT_NEED(nm, isf)                                                               \
  idx = bt_find_table(n, names, #nm);                                                  \
  if (idx < 0) { snprintf(err, errlen, "missing table '%s'", #nm); return -1; }        \
  if ((is_float[idx] != 0) != (isf)) { snprintf(err, errlen, "table '%s' has the wrong dtype", #nm); return -1; }
#define X(nm) BT_NEED(nm, 0) if (counts[idx] != 1) { snprintf(err, errlen, "scalar '%s' must have 1 element", #nm); return -1; } d->nm = *(const int*)host[idx];
  BT_INT_SCALARS(X)
#undef X
#define X(nm) BT_NEED(nm, 1) if (counts[idx] != 1) { snprintf(err, errlen, "scalar '%s' must have 1 element", #nm); return -1; } d->nm = *(const float*)host[idx];
  BT_FLT_SCALARS(X)
#undef X
#define X(nm) BT_NEED(nm, 0) d->nm = (const int*)bound[idx];
  BT_INT_TABLES(X)
#undef X
#define X(nm) BT_NEED(nm, 1) d->nm = (const float*)bound[idx];
  BT_FLT_TABLES(X)
#undef X
#undef BT_NEED
  // size checks for the tables whose extents the kernels derive from the scalars
  struct { const char* nm; int64_t want; } chk[] = {
      {"body_parentid", d->nbody}, {"body_rec", 12LL * d->nbody}, {"bl_rec", 16LL * d->nbody}, {"jnt_rec", 12LL * d->njnt},
      {"jnt_type", d->njnt}, {"qpos0", d->nq}, {"dof_rec", 8LL * d->nv}, {"dof_irec", d->nv}, {"dof_wgrp", d->nv},
      {"hpass_desc", 256LL * d->nhpass}, {"apass_desc", 32LL * d->napass}, {"cmp_adr", d->nbanc + 1LL},
      {"dof_range", 2LL * d->nv}, {"dof_solimp", 5LL * d->nv}, {"cbcon_adr", d->ncb + 1LL}, {"wgrp_adr", d->nwgrp + 1LL},
      {"act_rec", 16LL * (d->nu > 0 ? d->nu : 1)},
      // clip tables: n_clips clips of clip_len frames stacked on the leading axis (preprocess.py:254-258)
      {"clip_position", 3LL * d->n_clips * d->clip_len * d->n_animals}, {"clip_quaternion", 4LL * d->n_clips * d->clip_len * d->n_animals},
      {"clip_joints", (int64_t)d->n_clips * d->clip_len * d->clip_nj}, {"clip_body_positions", 3LL * d->n_clips * d->clip_len * d->nbody},
      {"clip_angular_velocity", 3LL * d->n_clips * d->clip_len * d->n_animals}, {"joint_idxs", d->n_joint_idxs}, {"body_idxs", d->n_body_idxs},
      {"animal_rec", 8LL * d->n_animals}, {"jidx_adr", d->n_animals + 1LL}, {"bidx_adr", d->n_animals + 1LL}, {"eidx_adr", d->n_animals + 1LL},
  };
  for (size_t k = 0; k < sizeof(chk) / sizeof(chk[0]); k++) {
    idx = bt_find_table(n, names, chk[k].nm);
    if (idx < 0 || counts[idx] != chk[k].want) {
      snprintf(err, errlen, "table '%s' has %lld elements, expected %lld", chk[k].nm, idx < 0 ? -1LL : (long long)counts[idx],
               (long long)chk[k].want);
      return -1;
    }
  }
  if (d->ncon > 0) {
    idx = bt_find_table(n, names, "con_g1");
    if (counts[idx] != d->ncon) { snprintf(err, errlen, "con_* tables must have ncon rows"); return -1; }
    idx = bt_find_table(n, names, "con_xref");
    if (counts[idx] != d->ncon) { snprintf(err, errlen, "con_xref must have ncon rows"); return -1; }
  }
  idx = bt_find_table(n, names, "sh_tab");
  if (counts[idx] < d->sh_stage_floats || counts[idx] < d->sho_body_rec + 12LL * d->nbody || counts[idx] < d->sho_bl_rec + 16LL * d->nbody ||
      counts[idx] < d->sho_jnt_rec + 12LL * d->njnt || d->sho_body_rec < 0 || d->sho_bl_rec < 0 || d->sho_jnt_rec < 0 ||
      d->sho_wrap_rec < 0 || d->sho_dofact_rec < 0 || d->sho_wrap_rec >= counts[idx] || d->sho_dofact_rec >= counts[idx] ||
      ((d->sho_body_rec | d->sho_bl_rec | d->sho_jnt_rec | d->sho_wrap_rec) & 3) || (d->sho_dofact_rec & 1)) {
    snprintf(err, errlen, "sh_tab is shorter than its layout");
    return -1;
  }
  if (d->clip_len < d->ref_len || d->nv <= 0 || d->nbody <= 1 || d->n_animals < 1 || d->n_clips < 1) { snprintf(err, errlen, "degenerate model"); return -1; }
  if (d->obs_size + 3 > d->smem_floats - d->o_crb) { snprintf(err, errlen, "observation row does not fit the staging region"); return -1; }
  return 0;
}

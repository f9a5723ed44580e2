/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C, dense formulation) of the physics half of the reference hot path:
 *   /root/reference/envs/fruitfly.py:500  self.pipeline_step(data0, action)
 *     -> brax.envs.base.PipelineEnv.pipeline_step  (scan x n_frames)
 *     -> brax.mjx.pipeline.step -> mujoco.mjx.step = forward + euler
 * The arithmetic lives in third-party, un-vendored, unpinned packages (mujoco-mjx ~3.2.x, brax ~0.10.x;
 * SURVEY.md F2), none importable in this image: **parity unpinned**.  The functions below restate the
 * published MJX algorithm, in the *dense* form the reference forces with
 * `mj_model.opt.jacobian = 0` (fruitfly.py:78): dense CRB mass matrix, dense Cholesky, dense efc_J.
 * The CUDA product uses a different (tree-sparse, on-chip) formulation, so agreement is meaningful.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Build: see oracle/Makefile (REAL = double by default, -DREAL_FLOAT for the fp32 build).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifdef REAL_FLOAT
typedef float real;
#define RSQRT sqrtf
#define RFABS fabsf
#else
typedef double real;
#define RSQRT sqrt
#define RFABS fabs
#endif

#define mjMINVAL ((real)1e-15)
#define mjMINIMP ((real)0.0001)
#define mjMAXIMP ((real)0.9999)

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_ELLIPSOID = 4 };
enum { CONE_PYRAMIDAL = 0, CONE_ELLIPTIC = 1 };

/* ---- model: field names follow mjModel; filled from Python (oracle/oracle.py) ---- */
typedef struct {
  int nq, nv, nu, na, nbody, njnt, ngeom, ntendon, npair, ncon, nefc_max, nlimit;
  int cone, iterations, ls_iterations;
  double timestep, gravity[3], density, viscosity, impratio, tolerance, ls_tolerance, meaninertia;
  const int *body_parentid, *body_rootid, *body_jntadr, *body_jntnum, *body_dofadr, *body_dofnum;
  const int *jnt_type, *jnt_qposadr, *jnt_dofadr, *jnt_bodyid, *jnt_limited;
  const int *dof_bodyid, *dof_jntid, *dof_parentid;
  const int *geom_type, *geom_bodyid;
  const int *pair_geom, *pair_condim, *pair_ncon;
  const int *tendon_adr, *tendon_num, *wrap_jntid;
  const int *actuator_trntype, *actuator_trnid, *actuator_dyntype, *actuator_gaintype, *actuator_biastype,
      *actuator_ctrllimited, *actuator_forcelimited, *actuator_actadr;
  const double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_invweight0;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_stiffness, *jnt_solref, *jnt_solimp, *jnt_margin;
  const double *qpos0, *qpos_spring, *dof_armature, *dof_damping, *dof_invweight0;
  const double *geom_pos, *geom_quat, *geom_size;
  const double *pair_friction, *pair_solref, *pair_solimp, *pair_margin, *pair_gap;
  const double *wrap_coef;
  const double *actuator_gear, *actuator_gainprm, *actuator_biasprm, *actuator_dynprm, *actuator_ctrlrange,
      *actuator_forcerange;
} OModel;

/* ---- data: every intermediate is kept so tests can inspect it ---- */
typedef struct {
  real *qpos, *qvel, *act, *ctrl, *qacc_warmstart, *time;
  real *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *geom_xpos, *geom_xmat;
  real *subtree_com, *cinert, *cdof, *crb, *qM, *qLD;
  real *cvel, *cdof_dot, *qfrc_bias, *qfrc_passive;
  real *actuator_length, *actuator_velocity, *actuator_force, *act_dot, *qfrc_actuator, *actuator_moment;
  real *qfrc_smooth, *qacc_smooth, *qacc, *qfrc_constraint;
  real *con_dist, *con_pos, *con_frame, *con_friction, *con_solref, *con_solimp, *con_includemargin;
  int *con_geom, *con_dim;
  real *efc_J, *efc_D, *efc_aref, *efc_pos, *efc_force;
  int *efc_type; /* 0 limit, 1 frictionless/pyramidal contact row, 2 elliptic normal row, 3 elliptic friction row */
  int *efc_id;   /* contact id for contact rows */
  int *nefc, *solver_niter;
  real *scratch; /* >= 16*nv + 8*nefc_max + nv*nv */
} OData;

/* ------------------------------------------------------------------------------------------ */
static inline void cross3(const real *a, const real *b, real *o) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline real dot3(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline real normalize3(real *a) {
  real n = RSQRT(dot3(a, a));
  if (n < mjMINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }   /* never hit by the selected models */
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
static inline void quat_mul(const real *a, const real *b, real *o) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
static inline void quat_normalize(real *q) {
  real n = RSQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
/* rotate vector by quaternion (mjx math.rotate) */
static inline void rotate(const real *v, const real *q, real *o) {
  real s = q[0];
  const real *u = q + 1;
  real uv = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(u, v, c);
  real r0 = 2 * uv * u[0] + (s * s - uu) * v[0] + 2 * s * c[0];
  real r1 = 2 * uv * u[1] + (s * s - uu) * v[1] + 2 * s * c[1];
  real r2 = 2 * uv * u[2] + (s * s - uu) * v[2] + 2 * s * c[2];
  o[0] = r0; o[1] = r1; o[2] = r2;
}
static inline void quat_to_mat(const real *q, real *m) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void axis_angle_to_quat(const real *axis, real angle, real *q) {
  real s = sin(angle * (real)0.5), c = cos(angle * (real)0.5);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
static inline void mat_vec3(const real *m, const real *v, real *o) {
  real a = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  real b = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  real c = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = a; o[1] = b; o[2] = c;
}
static inline void matT_vec3(const real *m, const real *v, real *o) {
  real a = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  real b = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  real c = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  o[0] = a; o[1] = b; o[2] = c;
}
/* 10-number spatial inertia times 6-vector [ang; lin]  (mjx math.inert_mul) */
static inline void inert_mul(const real *i, const real *v, real *o) {
  real ang[3], c[3];
  ang[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2];
  ang[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2];
  ang[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2];
  cross3(i + 6, v + 3, c);
  real o0 = ang[0] + c[0], o1 = ang[1] + c[1], o2 = ang[2] + c[2];
  cross3(i + 6, v, c);
  real o3 = i[9] * v[3] - c[0], o4 = i[9] * v[4] - c[1], o5 = i[9] * v[5] - c[2];
  o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = o5;
}
/* mjx math.motion_cross: u x v for motion vectors */
static inline void motion_cross(const real *u, const real *v, real *o) {
  real a[3], b[3], c[3];
  cross3(u, v, a);
  cross3(u + 3, v, b);
  cross3(u, v + 3, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
/* mjx math.motion_cross_force: v x* f */
static inline void motion_cross_force(const real *v, const real *f, real *o) {
  real a[3], b[3], c[3];
  cross3(v, f, a);
  cross3(v + 3, f + 3, b);
  cross3(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}

/* ---------------------------------- smooth.kinematics ------------------------------------- */
void o_kinematics(const OModel *m, OData *d) {
  for (int b = 0; b < m->nbody; b++) {
    real pos[3], quat[4];
    if (b == 0) {
      pos[0] = pos[1] = pos[2] = 0; quat[0] = 1; quat[1] = quat[2] = quat[3] = 0;
    } else {
      int p = m->body_parentid[b];
      real bp[3] = {(real)m->body_pos[3 * b], (real)m->body_pos[3 * b + 1], (real)m->body_pos[3 * b + 2]};
      real bq[4] = {(real)m->body_quat[4 * b], (real)m->body_quat[4 * b + 1], (real)m->body_quat[4 * b + 2], (real)m->body_quat[4 * b + 3]};
      rotate(bp, d->xquat + 4 * p, pos);
      for (int k = 0; k < 3; k++) pos[k] += d->xpos[3 * p + k];
      quat_mul(d->xquat + 4 * p, bq, quat);
    }
    for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
      int j = m->body_jntadr[b] + jj, qa = m->jnt_qposadr[j];
      real jp[3] = {(real)m->jnt_pos[3 * j], (real)m->jnt_pos[3 * j + 1], (real)m->jnt_pos[3 * j + 2]};
      real ja[3] = {(real)m->jnt_axis[3 * j], (real)m->jnt_axis[3 * j + 1], (real)m->jnt_axis[3 * j + 2]};
      if (m->jnt_type[j] == JNT_FREE) {
        for (int k = 0; k < 3; k++) { d->xanchor[3 * j + k] = d->qpos[qa + k]; pos[k] = d->qpos[qa + k]; }
        d->xaxis[3 * j] = 0; d->xaxis[3 * j + 1] = 0; d->xaxis[3 * j + 2] = 1;
        for (int k = 0; k < 4; k++) quat[k] = d->qpos[qa + 3 + k];
        quat_normalize(quat);
        for (int k = 0; k < 4; k++) d->qpos[qa + 3 + k] = quat[k]; /* kinematics writes back normalised quats */
      } else { /* hinge */
        real anchor[3], axis[3], qloc[4], q2[4], r[3];
        rotate(jp, quat, anchor);
        for (int k = 0; k < 3; k++) anchor[k] += pos[k];
        rotate(ja, quat, axis);
        for (int k = 0; k < 3; k++) { d->xanchor[3 * j + k] = anchor[k]; d->xaxis[3 * j + k] = axis[k]; }
        axis_angle_to_quat(ja, d->qpos[qa] - (real)m->qpos0[qa], qloc);
        quat_mul(quat, qloc, q2);
        for (int k = 0; k < 4; k++) quat[k] = q2[k];
        rotate(jp, quat, r);
        for (int k = 0; k < 3; k++) pos[k] = anchor[k] - r[k];
      }
    }
    for (int k = 0; k < 3; k++) d->xpos[3 * b + k] = pos[k];
    for (int k = 0; k < 4; k++) d->xquat[4 * b + k] = quat[k];
    quat_to_mat(quat, d->xmat + 9 * b);
    real ip[3] = {(real)m->body_ipos[3 * b], (real)m->body_ipos[3 * b + 1], (real)m->body_ipos[3 * b + 2]};
    real iq[4] = {(real)m->body_iquat[4 * b], (real)m->body_iquat[4 * b + 1], (real)m->body_iquat[4 * b + 2], (real)m->body_iquat[4 * b + 3]};
    real r[3], q2[4];
    rotate(ip, quat, r);
    for (int k = 0; k < 3; k++) d->xipos[3 * b + k] = pos[k] + r[k];
    quat_mul(quat, iq, q2);
    quat_to_mat(q2, d->ximat + 9 * b);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    real gp[3] = {(real)m->geom_pos[3 * g], (real)m->geom_pos[3 * g + 1], (real)m->geom_pos[3 * g + 2]};
    real gq[4] = {(real)m->geom_quat[4 * g], (real)m->geom_quat[4 * g + 1], (real)m->geom_quat[4 * g + 2], (real)m->geom_quat[4 * g + 3]};
    real r[3], q2[4];
    rotate(gp, d->xquat + 4 * b, r);
    for (int k = 0; k < 3; k++) d->geom_xpos[3 * g + k] = d->xpos[3 * b + k] + r[k];
    quat_mul(d->xquat + 4 * b, gq, q2);
    quat_to_mat(q2, d->geom_xmat + 9 * g);
  }
}

/* ---------------------------------- smooth.com_pos ---------------------------------------- */
void o_com_pos(const OModel *m, OData *d) {
  int nb = m->nbody;
  real *mass = d->scratch; /* nb */
  for (int b = 0; b < nb; b++) {
    mass[b] = (real)m->body_mass[b];
    for (int k = 0; k < 3; k++) d->subtree_com[3 * b + k] = d->xipos[3 * b + k] * mass[b];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    for (int k = 0; k < 3; k++) d->subtree_com[3 * p + k] += d->subtree_com[3 * b + k];
    mass[p] += mass[b];
  }
  for (int b = 0; b < nb; b++) {
    for (int k = 0; k < 3; k++) {
      if (mass[b] > 0) d->subtree_com[3 * b + k] /= mass[b];
      else d->subtree_com[3 * b + k] = d->xipos[3 * b + k];
    }
  }
  /* cinert: body inertia about the subtree com of the body's kinematic root, world axes */
  for (int b = 0; b < nb; b++) {
    const real *R = d->ximat + 9 * b;
    real mb = (real)m->body_mass[b];
    real off[3];
    for (int k = 0; k < 3; k++) off[k] = d->xipos[3 * b + k] - d->subtree_com[3 * m->body_rootid[b] + k];
    real I[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += R[3 * r + k] * (real)m->body_inertia[3 * b + k] * R[3 * c + k];
        I[3 * r + c] = s;
      }
    real oo = dot3(off, off);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) I[3 * r + c] += mb * ((r == c ? oo : 0) - off[r] * off[c]);
    real *ci = d->cinert + 10 * b;
    ci[0] = I[0]; ci[1] = I[4]; ci[2] = I[8]; ci[3] = I[1]; ci[4] = I[2]; ci[5] = I[5];
    ci[6] = off[0] * mb; ci[7] = off[1] * mb; ci[8] = off[2] * mb; ci[9] = mb;
  }
  /* cdof */
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    real off[3];
    for (int k = 0; k < 3; k++) off[k] = d->subtree_com[3 * m->body_rootid[b] + k] - d->xanchor[3 * j + k];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        real *c = d->cdof + 6 * (da + k);
        for (int i = 0; i < 6; i++) c[i] = 0;
        c[3 + k] = 1;
      }
      const real *R = d->xmat + 9 * b;
      for (int k = 0; k < 3; k++) {
        real *c = d->cdof + 6 * (da + 3 + k);
        real ax[3] = {R[k], R[3 + k], R[6 + k]};
        c[0] = ax[0]; c[1] = ax[1]; c[2] = ax[2];
        cross3(ax, off, c + 3);
      }
    } else {
      real *c = d->cdof + 6 * da;
      const real *ax = d->xaxis + 3 * j;
      c[0] = ax[0]; c[1] = ax[1]; c[2] = ax[2];
      cross3(ax, off, c + 3);
    }
  }
}

/* ---------------------------------- smooth.crb + factor_m (dense) ------------------------- */
void o_crb(const OModel *m, OData *d) {
  int nb = m->nbody, nv = m->nv;
  memcpy(d->crb, d->cinert, sizeof(real) * 10 * nb);
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * b + k];
  }
  for (int k = 0; k < 10; k++) d->crb[k] = 0;
  memset(d->qM, 0, sizeof(real) * nv * nv);
  for (int i = 0; i < nv; i++) {
    real buf[6];
    inert_mul(d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i, buf);
    int j = i;
    while (j >= 0) {
      real s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[6 * j + k] * buf[k];
      d->qM[i * nv + j] = s;
      d->qM[j * nv + i] = s;
      j = m->dof_parentid[j];
    }
    d->qM[i * nv + i] += (real)m->dof_armature[i];
  }
}

/* dense Cholesky A = L L^T (lower), returns 0 on success */
static int cholesky(const real *A, real *L, int n) {
  memset(L, 0, sizeof(real) * n * n);
  for (int j = 0; j < n; j++) {
    real s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
    if (s <= 0) return 1;
    real dj = RSQRT(s);
    L[j * n + j] = dj;
    for (int i = j + 1; i < n; i++) {
      real t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / dj;
    }
  }
  return 0;
}
static void cho_solve(const real *L, const real *b, real *x, int n) {
  for (int i = 0; i < n; i++) {
    real s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    real s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}
void o_factor_m(const OModel *m, OData *d) { cholesky(d->qM, d->qLD, m->nv); }
static void mul_m(const OModel *m, const OData *d, const real *v, real *o) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int j = 0; j < nv; j++) s += d->qM[i * nv + j] * v[j];
    o[i] = s;
  }
}

/* ---------------------------------- smooth.com_vel ---------------------------------------- */
void o_com_vel(const OModel *m, OData *d) {
  for (int b = 0; b < m->nbody; b++) {
    real cvel[6] = {0, 0, 0, 0, 0, 0};
    if (b > 0) memcpy(cvel, d->cvel + 6 * m->body_parentid[b], sizeof(cvel));
    for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
      int j = m->body_jntadr[b] + jj, da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JNT_FREE) {
        for (int k = 0; k < 3; k++)
          for (int i = 0; i < 6; i++) cvel[i] += d->cdof[6 * (da + k) + i] * d->qvel[da + k];
        for (int k = 0; k < 3; k++) {
          for (int i = 0; i < 6; i++) d->cdof_dot[6 * (da + k) + i] = 0;
          motion_cross(cvel, d->cdof + 6 * (da + 3 + k), d->cdof_dot + 6 * (da + 3 + k));
        }
        for (int k = 3; k < 6; k++)
          for (int i = 0; i < 6; i++) cvel[i] += d->cdof[6 * (da + k) + i] * d->qvel[da + k];
      } else {
        motion_cross(cvel, d->cdof + 6 * da, d->cdof_dot + 6 * da);
        for (int i = 0; i < 6; i++) cvel[i] += d->cdof[6 * da + i] * d->qvel[da];
      }
    }
    memcpy(d->cvel + 6 * b, cvel, sizeof(cvel));
  }
}

/* ---------------------------------- smooth.rne -------------------------------------------- */
void o_rne(const OModel *m, OData *d) {
  int nb = m->nbody, nv = m->nv;
  real *cacc = d->scratch;          /* nb*6 */
  real *cfrc = d->scratch + 6 * nb; /* nb*6 */
  for (int b = 0; b < nb; b++) {
    real *a = cacc + 6 * b;
    if (b == 0) {
      a[0] = a[1] = a[2] = 0;
      for (int k = 0; k < 3; k++) a[3 + k] = -(real)m->gravity[k];
    } else {
      memcpy(a, cacc + 6 * m->body_parentid[b], sizeof(real) * 6);
    }
    for (int dd = 0; dd < m->body_dofnum[b]; dd++) {
      int i = m->body_dofadr[b] + dd;
      for (int k = 0; k < 6; k++) a[k] += d->cdof_dot[6 * i + k] * d->qvel[i];
    }
  }
  for (int b = 0; b < nb; b++) {
    real t1[6], t2[6], t3[6];
    inert_mul(d->cinert + 10 * b, cacc + 6 * b, t1);
    inert_mul(d->cinert + 10 * b, d->cvel + 6 * b, t2);
    motion_cross_force(d->cvel + 6 * b, t2, t3);
    for (int k = 0; k < 6; k++) cfrc[6 * b + k] = t1[k] + t3[k];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    for (int k = 0; k < 6; k++) cfrc[6 * p + k] += cfrc[6 * b + k];
  }
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int k = 0; k < 6; k++) s += d->cdof[6 * i + k] * cfrc[6 * m->dof_bodyid[i] + k];
    d->qfrc_bias[i] = s;
  }
}

/* translational + rotational Jacobian of a world point attached to `body` (mjx support.jac) */
static void jac_point(const OModel *m, const OData *d, const real *point, int body, real *jacp, real *jacr) {
  int nv = m->nv;
  memset(jacp, 0, sizeof(real) * 3 * nv);
  if (jacr) memset(jacr, 0, sizeof(real) * 3 * nv);
  real off[3];
  for (int k = 0; k < 3; k++) off[k] = point[k] - d->subtree_com[3 * m->body_rootid[body] + k];
  /* last dof on the chain root->body */
  int b = body;
  while (b > 0 && m->body_dofnum[b] == 0) b = m->body_parentid[b];
  if (b == 0) return;
  int i = m->body_dofadr[b] + m->body_dofnum[b] - 1;
  while (i >= 0) {
    const real *c = d->cdof + 6 * i;
    real cr[3];
    cross3(c, off, cr);
    for (int k = 0; k < 3; k++) {
      jacp[k * nv + i] = c[3 + k] + cr[k];
      if (jacr) jacr[k * nv + i] = c[k];
    }
    i = m->dof_parentid[i];
  }
}

/* ---------------------------------- passive.passive --------------------------------------- */
void o_passive(const OModel *m, OData *d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = 0;
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] != JNT_HINGE) continue;
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    d->qfrc_passive[da] -= (real)m->jnt_stiffness[j] * (d->qpos[qa] - (real)m->qpos_spring[qa]);
  }
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] -= (real)m->dof_damping[i] * d->qvel[i];
  if (m->density > 0 || m->viscosity > 0) {
    /* inertia-box fluid model (mjx passive._inertia_box_fluid_model) */
    real *jacp = d->scratch, *jacr = d->scratch + 3 * nv;
    for (int b = 1; b < m->nbody; b++) {
      real mass = (real)m->body_mass[b];
      if (mass <= 0) continue;
      const real *in = m->body_inertia ? NULL : NULL;
      (void)in;
      real I0 = (real)m->body_inertia[3 * b], I1 = (real)m->body_inertia[3 * b + 1], I2 = (real)m->body_inertia[3 * b + 2];
      real box[3] = {I1 + I2 - I0, I0 + I2 - I1, I0 + I1 - I2};
      for (int k = 0; k < 3; k++) {
        real v = box[k] < (real)1e-12 ? (real)1e-12 : box[k];
        box[k] = RSQRT(6 * v / (mass > (real)1e-12 ? mass : (real)1e-12));
      }
      /* local 6D velocity at xipos in inertial-frame axes */
      const real *cv = d->cvel + 6 * b;
      const real *R = d->ximat + 9 * b;
      real off[3], lin[3], c[3], lang[3], llin[3];
      for (int k = 0; k < 3; k++) off[k] = d->xipos[3 * b + k] - d->subtree_com[3 * m->body_rootid[b] + k];
      cross3(off, cv, c);
      for (int k = 0; k < 3; k++) lin[k] = cv[3 + k] - c[k];
      matT_vec3(R, cv, lang);
      matT_vec3(R, lin, llin);
      real lfrc[6] = {0, 0, 0, 0, 0, 0};
      if (m->viscosity > 0) {
        real diam = (box[0] + box[1] + box[2]) / 3;
        for (int k = 0; k < 3; k++) {
          lfrc[k] = -lang[k] * (real)M_PI * diam * diam * diam * (real)m->viscosity;
          lfrc[3 + k] = -llin[k] * 3 * (real)M_PI * diam * (real)m->viscosity;
        }
      }
      if (m->density > 0) {
        real sv[3] = {box[1] * box[2], box[0] * box[2], box[0] * box[1]};
        real b4[3] = {box[0] * box[0] * box[0] * box[0], box[1] * box[1] * box[1] * box[1], box[2] * box[2] * box[2] * box[2]};
        real sa[3] = {box[0] * (b4[1] + b4[2]), box[1] * (b4[0] + b4[2]), box[2] * (b4[0] + b4[1])};
        for (int k = 0; k < 3; k++) {
          lfrc[3 + k] -= (real)0.5 * (real)m->density * sv[k] * RFABS(llin[k]) * llin[k];
          lfrc[k] -= (real)m->density * sa[k] * RFABS(lang[k]) * lang[k] / 64;
        }
      }
      real torque[3], force[3];
      mat_vec3(R, lfrc, torque);
      mat_vec3(R, lfrc + 3, force);
      jac_point(m, d, d->xipos + 3 * b, b, jacp, jacr);
      for (int i = 0; i < nv; i++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += jacp[k * nv + i] * force[k] + jacr[k * nv + i] * torque[k];
        d->qfrc_passive[i] += s;
      }
    }
  }
}

/* ---------------------------------- transmission + fwd_actuation -------------------------- */
void o_actuation(const OModel *m, OData *d) {
  int nv = m->nv, nu = m->nu;
  memset(d->actuator_moment, 0, sizeof(real) * nu * nv);
  for (int i = 0; i < nv; i++) d->qfrc_actuator[i] = 0;
  for (int u = 0; u < nu; u++) {
    real gear = (real)m->actuator_gear[u], len = 0;
    real *mom = d->actuator_moment + u * nv;
    if (m->actuator_trntype[u] == 0) {
      int j = m->actuator_trnid[u];
      len = d->qpos[m->jnt_qposadr[j]] * gear;
      mom[m->jnt_dofadr[j]] = gear;
    } else {
      int t = m->actuator_trnid[u];
      for (int w = 0; w < m->tendon_num[t]; w++) {
        int j = m->wrap_jntid[m->tendon_adr[t] + w];
        real coef = (real)m->wrap_coef[m->tendon_adr[t] + w];
        len += coef * d->qpos[m->jnt_qposadr[j]];
        mom[m->jnt_dofadr[j]] = coef * gear;
      }
      len *= gear;
    }
    real vel = 0;
    for (int i = 0; i < nv; i++) vel += mom[i] * d->qvel[i];
    d->actuator_length[u] = len;
    d->actuator_velocity[u] = vel;
    real ctrl = d->ctrl[u];
    if (m->actuator_ctrllimited[u]) {
      real lo = (real)m->actuator_ctrlrange[2 * u], hi = (real)m->actuator_ctrlrange[2 * u + 1];
      ctrl = ctrl < lo ? lo : (ctrl > hi ? hi : ctrl);
    }
    real ctrl_act = ctrl;
    int aa = m->actuator_actadr[u];
    if (aa >= 0) { /* filter */
      real tau = (real)m->actuator_dynprm[3 * u];
      if (tau < mjMINVAL) tau = mjMINVAL;
      d->act_dot[aa] = (ctrl - d->act[aa]) / tau;
      ctrl_act = d->act[aa];
    }
    const double *gp = m->actuator_gainprm + 3 * u, *bp = m->actuator_biasprm + 3 * u;
    real gain = (real)gp[0];
    if (m->actuator_gaintype[u] == 1) gain += (real)gp[1] * len + (real)gp[2] * vel;
    real bias = 0;
    if (m->actuator_biastype[u] == 1) bias = (real)bp[0] + (real)bp[1] * len + (real)bp[2] * vel;
    real force = gain * ctrl_act + bias;
    if (m->actuator_forcelimited[u]) {
      real lo = (real)m->actuator_forcerange[2 * u], hi = (real)m->actuator_forcerange[2 * u + 1];
      force = force < lo ? lo : (force > hi ? hi : force);
    }
    d->actuator_force[u] = force;
    for (int i = 0; i < nv; i++) d->qfrc_actuator[i] += mom[i] * force;
  }
}

/* ---------------------------------- collision (static pair list) -------------------------- */
/* mjx math.make_frame */
static void make_frame(const real *a_in, real *frame) {
  real a[3] = {a_in[0], a_in[1], a_in[2]};
  normalize3(a);
  real b[3] = {0, 0, 0};
  if (a[1] > (real)-0.5 && a[1] < (real)0.5) b[1] = 1; else b[2] = 1;
  real ab = dot3(a, b);
  for (int k = 0; k < 3; k++) b[k] -= a[k] * ab;
  normalize3(b);
  real c[3];
  cross3(a, b, c);
  for (int k = 0; k < 3; k++) { frame[k] = a[k]; frame[3 + k] = b[k]; frame[6 + k] = c[k]; }
}

static void closest_segment_to_segment(const real *a0, const real *a1, const real *b0, const real *b1, real *pa, real *pb) {
  /* mjx math.closest_segment_to_segment_points */
  real dir_a[3], dir_b[3], len_a, len_b;
  for (int k = 0; k < 3; k++) { dir_a[k] = a1[k] - a0[k]; dir_b[k] = b1[k] - b0[k]; }
  len_a = RSQRT(dot3(dir_a, dir_a)); len_b = RSQRT(dot3(dir_b, dir_b));
  for (int k = 0; k < 3; k++) { dir_a[k] /= (len_a > 0 ? len_a : 1); dir_b[k] /= (len_b > 0 ? len_b : 1); }
  real half_a = len_a * (real)0.5, half_b = len_b * (real)0.5;
  real a_mid[3], b_mid[3], trans[3];
  for (int k = 0; k < 3; k++) { a_mid[k] = a0[k] + dir_a[k] * half_a; b_mid[k] = b0[k] + dir_b[k] * half_b; trans[k] = a_mid[k] - b_mid[k]; }
  real dira_dot_dirb = dot3(dir_a, dir_b), dira_dot_trans = dot3(dir_a, trans), dirb_dot_trans = dot3(dir_b, trans);
  real denom = 1 - dira_dot_dirb * dira_dot_dirb;
  real orig_t_a = (-dira_dot_trans + dira_dot_dirb * dirb_dot_trans) / (denom + (real)1e-6);
  real orig_t_b = dirb_dot_trans + orig_t_a * dira_dot_dirb;
  real t_a = orig_t_a < -half_a ? -half_a : (orig_t_a > half_a ? half_a : orig_t_a);
  real t_b = orig_t_b < -half_b ? -half_b : (orig_t_b > half_b ? half_b : orig_t_b);
  real best_a[3], best_b[3];
  for (int k = 0; k < 3; k++) { best_a[k] = a_mid[k] + dir_a[k] * t_a; best_b[k] = b_mid[k] + dir_b[k] * t_b; }
  /* closest point on each segment to the other's candidate */
  real na[3], nb_[3], t;
  {
    real v[3]; for (int k = 0; k < 3; k++) v[k] = best_b[k] - a_mid[k];
    t = dot3(v, dir_a); t = t < -half_a ? -half_a : (t > half_a ? half_a : t);
    for (int k = 0; k < 3; k++) na[k] = a_mid[k] + dir_a[k] * t;
  }
  {
    real v[3]; for (int k = 0; k < 3; k++) v[k] = best_a[k] - b_mid[k];
    t = dot3(v, dir_b); t = t < -half_b ? -half_b : (t > half_b ? half_b : t);
    for (int k = 0; k < 3; k++) nb_[k] = b_mid[k] + dir_b[k] * t;
  }
  real d1[3], d2[3];
  for (int k = 0; k < 3; k++) { d1[k] = na[k] - best_b[k]; d2[k] = best_a[k] - nb_[k]; }
  if (dot3(d1, d1) < dot3(d2, d2)) { for (int k = 0; k < 3; k++) { pa[k] = na[k]; pb[k] = best_b[k]; } }
  else { for (int k = 0; k < 3; k++) { pa[k] = best_a[k]; pb[k] = nb_[k]; } }
}

void o_collision(const OModel *m, OData *d) {
  int c = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom[2 * p], g2 = m->pair_geom[2 * p + 1];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    const real *p1 = d->geom_xpos + 3 * g1, *R1 = d->geom_xmat + 9 * g1;
    const real *p2 = d->geom_xpos + 3 * g2, *R2 = d->geom_xmat + 9 * g2;
    real s2[3] = {(real)m->geom_size[3 * g2], (real)m->geom_size[3 * g2 + 1], (real)m->geom_size[3 * g2 + 2]};
    real s1[3] = {(real)m->geom_size[3 * g1], (real)m->geom_size[3 * g1 + 1], (real)m->geom_size[3 * g1 + 2]};
    int n0 = c;
    if (t1 == GEOM_PLANE && t2 == GEOM_CAPSULE) {
      real n[3] = {R1[2], R1[5], R1[8]}, axis[3] = {R2[2], R2[5], R2[8]};
      /* frame aligned with the capsule axis (mjx collision_primitive.plane_capsule) */
      real na = dot3(n, axis), b[3];
      for (int k = 0; k < 3; k++) b[k] = axis[k] - n[k] * na;
      real bn = RSQRT(dot3(b, b));
      real y[3] = {R1[1], R1[4], R1[7]}, z[3] = {R1[2], R1[5], R1[8]};
      if (bn < (real)0.5) {
        real ny = n[1];
        const real *alt = (ny > (real)-0.5 && ny < (real)0.5) ? y : z;
        for (int k = 0; k < 3; k++) b[k] = alt[k];
      } else {
        for (int k = 0; k < 3; k++) b[k] /= bn;
      }
      real fr[9];
      for (int k = 0; k < 3; k++) { fr[k] = n[k]; fr[3 + k] = b[k]; }
      cross3(n, b, fr + 6);
      for (int e = 0; e < 2; e++) {
        real sgn = e == 0 ? 1 : -1, sp[3];
        for (int k = 0; k < 3; k++) sp[k] = p2[k] + sgn * axis[k] * s2[1];
        real diff[3] = {sp[0] - p1[0], sp[1] - p1[1], sp[2] - p1[2]};
        real dist = dot3(diff, n) - s2[0];
        d->con_dist[c] = dist;
        for (int k = 0; k < 3; k++) d->con_pos[3 * c + k] = sp[k] - n[k] * (s2[0] + (real)0.5 * dist);
        memcpy(d->con_frame + 9 * c, fr, sizeof(fr));
        c++;
      }
    } else if (t1 == GEOM_PLANE && (t2 == GEOM_ELLIPSOID || t2 == GEOM_SPHERE)) {
      real n[3] = {R1[2], R1[5], R1[8]};
      real pos[3], dist;
      if (t2 == GEOM_SPHERE) {
        real diff[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        dist = dot3(diff, n) - s2[0];
        for (int k = 0; k < 3; k++) pos[k] = p2[k] - n[k] * (s2[0] + (real)0.5 * dist);
      } else {
        real ln[3], sup[3], w[3];
        matT_vec3(R2, n, ln);
        for (int k = 0; k < 3; k++) sup[k] = ln[k] * s2[k];
        normalize3(sup);
        for (int k = 0; k < 3; k++) sup[k] = -sup[k] * s2[k];
        mat_vec3(R2, sup, w);
        for (int k = 0; k < 3; k++) pos[k] = p2[k] + w[k];
        real diff[3] = {pos[0] - p1[0], pos[1] - p1[1], pos[2] - p1[2]};
        dist = dot3(diff, n);
        for (int k = 0; k < 3; k++) pos[k] -= n[k] * dist * (real)0.5;
      }
      d->con_dist[c] = dist;
      for (int k = 0; k < 3; k++) d->con_pos[3 * c + k] = pos[k];
      make_frame(n, d->con_frame + 9 * c);
      c++;
    } else if (t1 == GEOM_CAPSULE && t2 == GEOM_CAPSULE) {
      real ax1[3] = {R1[2], R1[5], R1[8]}, ax2[3] = {R2[2], R2[5], R2[8]};
      real a0[3], a1[3], b0[3], b1[3], pa[3], pb[3];
      for (int k = 0; k < 3; k++) {
        a0[k] = p1[k] - ax1[k] * s1[1]; a1[k] = p1[k] + ax1[k] * s1[1];
        b0[k] = p2[k] - ax2[k] * s2[1]; b1[k] = p2[k] + ax2[k] * s2[1];
      }
      closest_segment_to_segment(a0, a1, b0, b1, pa, pb);
      /* sphere-sphere */
      real n[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
      real len = RSQRT(dot3(n, n));
      if (len < mjMINVAL) { n[0] = 1; n[1] = 0; n[2] = 0; } else { n[0] /= len; n[1] /= len; n[2] /= len; }
      real dist = len - (s1[0] + s2[0]);
      d->con_dist[c] = dist;
      for (int k = 0; k < 3; k++) d->con_pos[3 * c + k] = pa[k] + n[k] * (s1[0] + (real)0.5 * dist);
      make_frame(n, d->con_frame + 9 * c);
      c++;
    }
    for (int i = n0; i < c; i++) {
      d->con_geom[2 * i] = g1; d->con_geom[2 * i + 1] = g2;
      d->con_dim[i] = m->pair_condim[p];
      d->con_includemargin[i] = (real)(m->pair_margin[p] - m->pair_gap[p]);
      for (int k = 0; k < 5; k++) d->con_friction[5 * i + k] = (real)m->pair_friction[5 * p + k];
      for (int k = 0; k < 2; k++) d->con_solref[2 * i + k] = (real)m->pair_solref[2 * p + k];
      for (int k = 0; k < 5; k++) d->con_solimp[5 * i + k] = (real)m->pair_solimp[5 * p + k];
    }
  }
}

/* ---------------------------------- constraint.make_constraint ---------------------------- */
static void kbi(const OModel *m, const real *solref, const real *solimp, real pos, real *k, real *b, real *imp) {
  real timeconst = solref[0], dampratio = solref[1];
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (timeconst < 2 * (real)m->timestep) timeconst = 2 * (real)m->timestep; /* refsafe */
  dmin = dmin < mjMINIMP ? mjMINIMP : (dmin > mjMAXIMP ? mjMAXIMP : dmin);
  dmax = dmax < mjMINIMP ? mjMINIMP : (dmax > mjMAXIMP ? mjMAXIMP : dmax);
  if (width < mjMINVAL) width = mjMINVAL;
  mid = mid < mjMINIMP ? mjMINIMP : (mid > mjMAXIMP ? mjMAXIMP : mid);
  if (power < 1) power = 1;
  *k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  *b = 2 / (dmax * timeconst);
  if (solref[0] <= 0) *k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0) *b = -solref[1] / dmax;
  real x = RFABS(pos) / width;
  real ia = (1 / pow(mid, power - 1)) * pow(x, power);
  real ib = 1 - (1 / pow(1 - mid, power - 1)) * pow(1 - x, power);
  real y = x < mid ? ia : ib;
  real im = dmin + y * (dmax - dmin);
  im = im < dmin ? dmin : (im > dmax ? dmax : im);
  if (x > 1) im = dmax;
  *imp = im;
}

static void add_row(const OModel *m, OData *d, const real *J, real pos_aref, real pos_imp, real invweight,
                    const real *solref, const real *solimp, int type, int id) {
  int nv = m->nv, r = *d->nefc;
  real k, b, imp;
  kbi(m, solref, solimp, pos_imp, &k, &b, &imp);
  real R = invweight * (1 - imp) / imp;
  if (R < mjMINVAL) R = mjMINVAL;
  real vel = 0;
  for (int i = 0; i < nv; i++) { d->efc_J[r * nv + i] = J[i]; vel += J[i] * d->qvel[i]; }
  d->efc_D[r] = 1 / R;
  d->efc_aref[r] = -b * vel - k * imp * pos_aref;
  d->efc_pos[r] = pos_aref;
  d->efc_type[r] = type;
  d->efc_id[r] = id;
  *d->nefc = r + 1;
}

void o_make_constraint(const OModel *m, OData *d) {
  int nv = m->nv;
  *d->nefc = 0;
  real *J = d->scratch;                /* nv */
  real *jacp1 = d->scratch + nv;       /* 3nv */
  real *jacp2 = d->scratch + 4 * nv;   /* 3nv */
  real *jacr1 = d->scratch + 7 * nv;   /* 3nv */
  real *jacr2 = d->scratch + 10 * nv;  /* 3nv */
  /* joint limits (inactive rows are dropped: a zero-J row never produces force) */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] != JNT_HINGE) continue;
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    real q = d->qpos[qa];
    real dmin = q - (real)m->jnt_range[2 * j], dmax = (real)m->jnt_range[2 * j + 1] - q;
    real pos = (dmin < dmax ? dmin : dmax) - (real)m->jnt_margin[j];
    if (!(pos < 0)) continue;
    for (int i = 0; i < nv; i++) J[i] = 0;
    J[da] = dmin < dmax ? 1 : -1;
    real sr[2] = {(real)m->jnt_solref[2 * j], (real)m->jnt_solref[2 * j + 1]};
    real si[5];
    for (int k = 0; k < 5; k++) si[k] = (real)m->jnt_solimp[5 * j + k];
    add_row(m, d, J, pos, pos, (real)m->dof_invweight0[da], sr, si, 0, j);
  }
  /* contacts */
  for (int c = 0; c < m->ncon; c++) {
    real dist = d->con_dist[c] - d->con_includemargin[c];
    if (!(dist < 0)) continue;
    int b1 = m->geom_bodyid[d->con_geom[2 * c]], b2 = m->geom_bodyid[d->con_geom[2 * c + 1]];
    jac_point(m, d, d->con_pos + 3 * c, b1, jacp1, jacr1);
    jac_point(m, d, d->con_pos + 3 * c, b2, jacp2, jacr2);
    const real *fr = d->con_frame + 9 * c;
    real t = (real)m->body_invweight0[2 * b1] + (real)m->body_invweight0[2 * b2];
    int dim = d->con_dim[c];
    /* diff_con[r][i] = frame[r] . (jacp2 - jacp1)[:, i] */
    real *dc = d->scratch + 13 * nv; /* 3nv */
    for (int r = 0; r < 3; r++)
      for (int i = 0; i < nv; i++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += fr[3 * r + k] * (jacp2[k * nv + i] - jacp1[k * nv + i]);
        dc[r * nv + i] = s;
      }
    const real *sr = d->con_solref + 2 * c, *si = d->con_solimp + 5 * c;
    if (dim == 1) {
      add_row(m, d, dc, dist, dist, t, sr, si, 1, c);
    } else if (m->cone == CONE_PYRAMIDAL) {
      if (dim != 3) continue; /* condim 4/6 unused */
      for (int a = 0; a < 2; a++) {
        real mu = d->con_friction[5 * c + a];
        for (int s = 0; s < 2; s++) {
          real f = s == 0 ? mu : -mu;
          for (int i = 0; i < nv; i++) J[i] = dc[i] + dc[(1 + a) * nv + i] * f;
          real iw = (t + f * f * t) * 2 * f * f / (real)m->impratio;
          add_row(m, d, J, dist, dist, iw, sr, si, 1, c);
        }
      }
    } else { /* elliptic, condim 3 */
      if (dim != 3) continue;
      real mu = d->con_friction[5 * c];
      add_row(m, d, dc, dist, dist, t, sr, si, 2, c);
      for (int a = 0; a < 2; a++) {
        /* friction rows: pos 0 for aref, impedance evaluated at the normal's penetration; R scaled so that
           R_fric = R_normal / impratio   (mj_makeImpedance: efc_R[j] = R[0]*... / impratio handled via invweight) */
        real iw = t / (real)m->impratio;
        (void)mu;
        add_row(m, d, dc + (1 + a) * nv, 0, dist, iw, sr, si, 3, c);
      }
    }
  }
}

/* ---------------------------------- solver (CG) ------------------------------------------- */
typedef struct {
  real *qacc, *Ma, *Jaref, *force, *qfrc_c, *grad, *Mgrad, *search, *mv, *jv;
  real gauss, cost, prev_cost;
} Ctx;

/* elliptic-cone bookkeeping for contact c starting at row r: returns zone, fills per-contact terms */
static real ell_mu(const OData *d, int c) { return d->con_friction[5 * c]; }

static void update_constraint(const OModel *m, OData *d, Ctx *x) {
  int nv = m->nv, ne = *d->nefc;
  real cost = 0;
  for (int r = 0; r < ne; r++) x->force[r] = 0;
  for (int r = 0; r < ne; r++) {
    int ty = d->efc_type[r];
    if (ty == 0 || ty == 1) {
      real ja = x->Jaref[r];
      if (ja < 0) { x->force[r] = -d->efc_D[r] * ja; cost += (real)0.5 * d->efc_D[r] * ja * ja; }
    } else if (ty == 2) {
      /* elliptic contact rows r, r+1, r+2 (MuJoCo PrimalUpdateConstraint, cone section) */
      int c = d->efc_id[r];
      real mu = ell_mu(d, c) ;
      real fri[2] = {d->con_friction[5 * c], d->con_friction[5 * c + 1]};
      real Dn = d->efc_D[r];
      /* dual-cone scaling: mu_regularised = mu * sqrt(R_fric / R_normal) => with D_f = D_n*impratio: */
      real mu0 = mu * RSQRT(Dn / d->efc_D[r + 1]);
      real u0 = x->Jaref[r] * mu0;
      real u1 = x->Jaref[r + 1] * fri[0], u2 = x->Jaref[r + 2] * fri[1];
      real N = u0, T = RSQRT(u1 * u1 + u2 * u2);
      real Dm = Dn / (mu0 * mu0 * (1 + mu0 * mu0));
      if (N >= mu0 * T || (T <= 0 && N >= 0)) {
        /* top zone: no force */
      } else if (mu0 * N + T <= 0 || (T <= 0 && N < 0)) {
        /* bottom zone: quadratic in all three rows */
        for (int k = 0; k < 3; k++) {
          real ja = x->Jaref[r + k];
          x->force[r + k] = -d->efc_D[r + k] * ja;
          cost += (real)0.5 * d->efc_D[r + k] * ja * ja;
        }
      } else {
        /* middle zone */
        real NmT = N - mu0 * T;
        cost += (real)0.5 * Dm * NmT * NmT;
        x->force[r] = -Dm * NmT * mu0;
        x->force[r + 1] = Dm * NmT * mu0 / T * u1 * fri[0];
        x->force[r + 2] = Dm * NmT * mu0 / T * u2 * fri[1];
      }
    }
  }
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int r = 0; r < ne; r++) s += d->efc_J[r * nv + i] * x->force[r];
    x->qfrc_c[i] = s;
  }
  real g = 0;
  for (int i = 0; i < nv; i++) g += (x->Ma[i] - d->qfrc_smooth[i]) * (x->qacc[i] - d->qacc_smooth[i]);
  x->gauss = (real)0.5 * g;
  x->prev_cost = x->cost;
  x->cost = cost + x->gauss;
}

static void update_gradient(const OModel *m, OData *d, Ctx *x) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) x->grad[i] = x->Ma[i] - d->qfrc_smooth[i] - x->qfrc_c[i];
  cho_solve(d->qLD, x->grad, x->Mgrad, nv);
}

static void ctx_init(const OModel *m, OData *d, Ctx *x, const real *qacc) {
  int nv = m->nv, ne = *d->nefc;
  for (int i = 0; i < nv; i++) x->qacc[i] = qacc[i];
  for (int r = 0; r < ne; r++) {
    real s = 0;
    for (int i = 0; i < nv; i++) s += d->efc_J[r * nv + i] * qacc[i];
    x->Jaref[r] = s - d->efc_aref[r];
  }
  mul_m(m, d, qacc, x->Ma);
  x->cost = INFINITY; x->prev_cost = 0;
  update_constraint(m, d, x);
}

typedef struct { real alpha, cost, d0, d1; } LSPoint;

static LSPoint ls_point(const OModel *m, const OData *d, const Ctx *x, real alpha, const real *quad_gauss, const real *quad) {
  int ne = *d->nefc;
  real q0 = quad_gauss[0], q1 = quad_gauss[1], q2 = quad_gauss[2];
  for (int r = 0; r < ne; r++) {
    int ty = d->efc_type[r];
    if (ty == 0 || ty == 1) {
      real v = x->Jaref[r] + alpha * x->jv[r];
      if (v < 0) { q0 += quad[3 * r]; q1 += quad[3 * r + 1]; q2 += quad[3 * r + 2]; }
    } else if (ty == 2) {
      int c = d->efc_id[r];
      real mu = d->con_friction[5 * c];
      real fri[2] = {d->con_friction[5 * c], d->con_friction[5 * c + 1]};
      real Dn = d->efc_D[r];
      real mu0 = mu * RSQRT(Dn / d->efc_D[r + 1]);
      real Dm = Dn / (mu0 * mu0 * (1 + mu0 * mu0));
      real u0 = x->Jaref[r] * mu0, v0 = x->jv[r] * mu0;
      real u1 = x->Jaref[r + 1] * fri[0], v1 = x->jv[r + 1] * fri[0];
      real u2 = x->Jaref[r + 2] * fri[1], v2 = x->jv[r + 2] * fri[1];
      real uu = u1 * u1 + u2 * u2, uv = u1 * v1 + u2 * v2, vv = v1 * v1 + v2 * v2;
      real N = u0 + alpha * v0;
      real Tsqr = uu + alpha * (2 * uv + alpha * vv);
      real T = RSQRT(Tsqr > 0 ? Tsqr : 0);
      if (N >= mu0 * T || (T <= 0 && N >= 0)) {
        /* top */
      } else if (mu0 * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int k = 0; k < 3; k++) { q0 += quad[3 * (r + k)]; q1 += quad[3 * (r + k) + 1]; q2 += quad[3 * (r + k) + 2]; }
      } else {
        /* middle zone: non-quadratic; add value and derivatives at alpha directly via a local expansion */
        real N1 = v0, T1 = (uv + alpha * vv) / T;
        real T2 = vv / T - (uv + alpha * vv) * T1 / (T * T);
        real NmT = N - mu0 * T;
        real c0 = (real)0.5 * Dm * NmT * NmT;
        real c1 = Dm * NmT * (N1 - mu0 * T1);
        real c2 = Dm * ((N1 - mu0 * T1) * (N1 - mu0 * T1) + NmT * (-mu0 * T2));
        /* express as quadratic around alpha: cost(a') = c0 + c1 (a'-a) + c2/2 (a'-a)^2 evaluated at a'=a */
        q0 += c0 - c1 * alpha + (real)0.5 * c2 * alpha * alpha;
        q1 += c1 - c2 * alpha;
        q2 += (real)0.5 * c2;
      }
    }
  }
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.d0 = 2 * alpha * q2 + q1;
  p.d1 = 2 * q2 + (q2 == 0 ? mjMINVAL : 0);
  return p;
}

static void linesearch(const OModel *m, OData *d, Ctx *x, real *quad) {
  int nv = m->nv, ne = *d->nefc;
  real sn = 0;
  for (int i = 0; i < nv; i++) sn += x->search[i] * x->search[i];
  real smag = RSQRT(sn) * (real)m->meaninertia * (nv > 1 ? nv : 1);
  real gtol = (real)m->tolerance * (real)m->ls_tolerance * smag;
  mul_m(m, d, x->search, x->mv);
  for (int r = 0; r < ne; r++) {
    real s = 0;
    for (int i = 0; i < nv; i++) s += d->efc_J[r * nv + i] * x->search[i];
    x->jv[r] = s;
  }
  real qg[3] = {x->gauss, 0, 0};
  for (int i = 0; i < nv; i++) {
    qg[1] += x->search[i] * (x->Ma[i] - d->qfrc_smooth[i]);
    qg[2] += (real)0.5 * x->search[i] * x->mv[i];
  }
  for (int r = 0; r < ne; r++) {
    quad[3 * r] = (real)0.5 * x->Jaref[r] * x->Jaref[r] * d->efc_D[r];
    quad[3 * r + 1] = x->jv[r] * x->Jaref[r] * d->efc_D[r];
    quad[3 * r + 2] = (real)0.5 * x->jv[r] * x->jv[r] * d->efc_D[r];
  }
  LSPoint p0 = ls_point(m, d, x, 0, qg, quad);
  LSPoint lo = ls_point(m, d, x, p0.alpha - p0.d0 / p0.d1, qg, quad);
  LSPoint hi;
  if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
  int swap = 1, it = 0;
  while (1) {
    int done = it >= m->ls_iterations;
    done |= !swap;
    done |= (lo.d0 < 0) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0) && (hi.d0 < gtol);
    if (done) break;
    LSPoint lo_next = ls_point(m, d, x, lo.alpha - lo.d0 / lo.d1, qg, quad);
    LSPoint hi_next = ls_point(m, d, x, hi.alpha - hi.d0 / hi.d1, qg, quad);
    LSPoint mid = ls_point(m, d, x, (real)0.5 * (lo.alpha + hi.alpha), qg, quad);
    int s_lo_next = (lo.d0 > 0) || (lo.d0 < lo_next.d0);
    if (s_lo_next) lo = lo_next;
    int s_lo_mid = (mid.d0 < 0) && (lo.d0 < mid.d0);
    if (s_lo_mid) lo = mid;
    int s_hi_next = (hi.d0 < 0) || (hi.d0 > hi_next.d0);
    if (s_hi_next) hi = hi_next;
    int s_hi_mid = (mid.d0 > 0) && (hi.d0 > mid.d0);
    if (s_hi_mid) hi = mid;
    swap = s_lo_next || s_lo_mid || s_hi_next || s_hi_mid;
    it++;
  }
  int improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  real alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (improved) {
    for (int i = 0; i < nv; i++) { x->qacc[i] += x->search[i] * alpha; x->Ma[i] += x->mv[i] * alpha; }
    for (int r = 0; r < ne; r++) x->Jaref[r] += x->jv[r] * alpha;
  }
}

void o_solve(const OModel *m, OData *d) {
  int nv = m->nv, ne = *d->nefc;
  if (ne == 0) {
    /* MJX has a static nefc > 0 for these models; with every row inactive the solver's answer is qacc_smooth
       up to rounding and warmstart = qacc. */
    for (int i = 0; i < nv; i++) { d->qacc[i] = d->qacc_smooth[i]; d->qacc_warmstart[i] = d->qacc_smooth[i]; d->qfrc_constraint[i] = 0; }
    *d->solver_niter = 0;
    return;
  }
  real *w = d->scratch + 16 * nv;
  Ctx x;
  x.qacc = w; w += nv; x.Ma = w; w += nv; x.qfrc_c = w; w += nv; x.grad = w; w += nv; x.Mgrad = w; w += nv;
  x.search = w; w += nv; x.mv = w; w += nv;
  real *pg = w; w += nv; real *pMg = w; w += nv;
  x.Jaref = w; w += m->nefc_max; x.force = w; w += m->nefc_max; x.jv = w; w += m->nefc_max;
  real *quad = w; w += 3 * m->nefc_max;
  /* warmstart selection */
  ctx_init(m, d, &x, d->qacc_warmstart);
  real cost_warm = x.cost;
  ctx_init(m, d, &x, d->qacc_smooth);
  real cost_smooth = x.cost;
  if (cost_warm < cost_smooth) ctx_init(m, d, &x, d->qacc_warmstart);
  update_gradient(m, d, &x);
  for (int i = 0; i < nv; i++) x.search[i] = -x.Mgrad[i];
  real scale = 1 / ((real)m->meaninertia * (nv > 1 ? nv : 1));
  int niter = 0;
  while (1) {
    if (m->iterations != 1) {
      real improvement = (x.prev_cost - x.cost) * scale;
      real gn = 0;
      for (int i = 0; i < nv; i++) gn += x.grad[i] * x.grad[i];
      real gradient = RSQRT(gn) * scale;
      if (niter >= m->iterations || improvement < (real)m->tolerance || gradient < (real)m->tolerance) break;
    } else if (niter >= 1) break;
    linesearch(m, d, &x, quad);
    for (int i = 0; i < nv; i++) { pg[i] = x.grad[i]; pMg[i] = x.Mgrad[i]; }
    update_constraint(m, d, &x);
    update_gradient(m, d, &x);
    real num = 0, den = 0;
    for (int i = 0; i < nv; i++) { num += x.grad[i] * (x.Mgrad[i] - pMg[i]); den += pg[i] * pMg[i]; }
    real beta = num / (den > mjMINVAL ? den : mjMINVAL);
    if (beta < 0) beta = 0;
    for (int i = 0; i < nv; i++) x.search[i] = -x.Mgrad[i] + beta * x.search[i];
    niter++;
  }
  *d->solver_niter = niter;
  for (int i = 0; i < nv; i++) { d->qacc[i] = x.qacc[i]; d->qacc_warmstart[i] = x.qacc[i]; d->qfrc_constraint[i] = x.qfrc_c[i]; }
  for (int r = 0; r < ne; r++) d->efc_force[r] = x.force[r];
}

/* ---------------------------------- forward / euler / step -------------------------------- */
void o_forward(const OModel *m, OData *d) {
  int nv = m->nv;
  o_kinematics(m, d);
  o_com_pos(m, d);
  o_crb(m, d);
  o_factor_m(m, d);
  o_collision(m, d);
  o_make_constraint(m, d);
  o_com_vel(m, d);
  o_passive(m, d);
  o_rne(m, d);
  o_actuation(m, d);
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  cho_solve(d->qLD, d->qfrc_smooth, d->qacc_smooth, nv);
  o_solve(m, d);
}

void o_euler(const OModel *m, OData *d) {
  int nv = m->nv;
  real h = (real)m->timestep;
  real *A = d->scratch + 16 * nv;  /* nv*nv */
  real *L = A + nv * nv;           /* nv*nv */
  real *rhs = L + nv * nv, *qacc = rhs + nv;
  memcpy(A, d->qM, sizeof(real) * nv * nv);
  for (int i = 0; i < nv; i++) A[i * nv + i] += h * (real)m->dof_damping[i];
  cholesky(A, L, nv);
  for (int i = 0; i < nv; i++) rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
  cho_solve(L, rhs, qacc, nv);
  for (int u = 0; u < m->nu; u++) {
    int aa = m->actuator_actadr[u];
    if (aa >= 0) d->act[aa] += d->act_dot[aa] * h;
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += qacc[i] * h;
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      real v[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
      real n = RSQRT(dot3(v, v));
      real ax[3] = {1, 0, 0}; /* mjx normalize_with_norm of a zero vector returns the zero vector; angle 0 */
      if (n > 0) { ax[0] = v[0] / n; ax[1] = v[1] / n; ax[2] = v[2] / n; } else { ax[0] = 0; }
      real qr[4], q2[4];
      axis_angle_to_quat(ax, h * n, qr);
      quat_mul(d->qpos + qa + 3, qr, q2);
      quat_normalize(q2);
      for (int k = 0; k < 4; k++) d->qpos[qa + 3 + k] = q2[k];
    } else {
      d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time[0] += h;
}

void o_step(const OModel *m, OData *d) {
  o_forward(m, d);
  o_euler(m, d);
}

int o_sizeof_real(void) { return (int)sizeof(real); }

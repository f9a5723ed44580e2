// Per-environment algorithm of the fused physics + tracking-reward step.
//
// One *lane group* of G lanes owns one environment (G = 32: a warp per environment on sm_100a; G = 1 is the
// host emulation used by the CPU-side tests of the table logic).  All per-environment state lives in a scratch
// block `s` (shared memory on the GPU) laid out by model.py::pack; model constants are read through BtDev.
//
// What this replaces (reference call path): /root/reference/envs/fruitfly.py:497-596 (env.step) ->
// brax PipelineEnv.pipeline_step -> mujoco.mjx.step x n_frames, plus
// custom_brax/custom_wrappers.py:54-80 and brax EpisodeWrapper.  The formulation is deliberately different
// from MJX's dense one (fruitfly.py:78): no mass matrix and no factor are ever formed -- the articulated-body
// recursion (aba_factor) yields what M^-1 v and M v need, both of which are two O(nv) chain sweeps; the
// constraint Jacobian is matrix-free (its chain sums are by-products of those sweeps); constraint rows are held
// in registers.  DESIGN.md section 2 describes the execution model.
#pragma once
#include "bt_math.h"
#include "bt_model.h"

template <int G>
struct BtLanes;
#ifdef __CUDACC__
template <>
struct BtLanes<32> {
  static BT_DEV void sync() { __syncwarp(); }
  static BT_DEV float allsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  // N sums at once by recursive halving: at every butterfly step a lane keeps one half of its values and trades the other
  // half with its partner, so the 5 steps cost ~N + log N shuffles instead of 5 N; the totals end up spread over the lanes
  // (value j in the lane whose bits select it) and are broadcast back.  Summation order differs from allsum() only in
  // association.
  template <int N>
  static BT_DEV void allsumN(float (&v)[N], int lane) {
    constexpr int h1 = (N + 1) / 2, h2 = (h1 + 1) / 2, h3 = (h2 + 1) / 2, h4 = (h3 + 1) / 2;
    float a[h1 > 0 ? h1 : 1];
    int src = 0;
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int i = 0; i < h1; i++) {
        const float up = h1 + i < N ? v[h1 + i] : 0.f;
        const float keep = hi ? up : v[i], send = hi ? v[i] : up;
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      src += hi ? h1 : 0;
    }
    if (h1 > 1) {
      const bool hi = lane & 8;
#pragma unroll
      for (int i = 0; i < h2; i++) {
        const float up = h2 + i < h1 ? a[h2 + i] : 0.f;
        const float keep = hi ? up : a[i], send = hi ? a[i] : up;
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    } else a[0] += __shfl_xor_sync(0xffffffffu, a[0], 8);
    if (h2 > 1) {
      const bool hi = lane & 4;
#pragma unroll
      for (int i = 0; i < h3; i++) {
        const float up = h3 + i < h2 ? a[h3 + i] : 0.f;
        const float keep = hi ? up : a[i], send = hi ? a[i] : up;
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    } else a[0] += __shfl_xor_sync(0xffffffffu, a[0], 4);
    if (h3 > 1) {
      const bool hi = lane & 2;
#pragma unroll
      for (int i = 0; i < h4; i++) {
        const float up = h4 + i < h3 ? a[h4 + i] : 0.f;
        const float keep = hi ? up : a[i], send = hi ? a[i] : up;
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
    } else a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
    static_assert(h4 == 1, "allsumN supports up to 16 values");
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
    // value j sits in the lanes whose (16, 8, 4, 2) bits spell its position in the halving tree
#pragma unroll
    for (int j = 0; j < N; j++) {
      int l = 0, r = j;
      if (r >= h1) { l |= 16; r -= h1; }
      if (h1 > 1 && r >= h2) { l |= 8; r -= h2; }
      if (h2 > 1 && r >= h3) { l |= 4; r -= h3; }
      if (h3 > 1 && r >= h4) { l |= 2; r -= h4; }
      v[j] = __shfl_sync(0xffffffffu, a[0], l);
    }
  }
  static BT_DEV int any(int p) { return __any_sync(0xffffffffu, p); }
  static BT_DEV int allmax(int v) { return __reduce_max_sync(0xffffffffu, v); }
  // CTA-wide phase alignment: the warps of a CTA own different environments but are kept in the same phase of the
  // program so that the instruction-fetch working set is one phase (I-cache: 6 KB L0 / 32 KB L1.5 vs a 200 KB kernel)
  static BT_DEV void cta_sync() { __syncthreads(); }
  static BT_DEV int cta_any(int p) { return __syncthreads_or(p); }
  // alignment of a subset of the CTA's warps only (named barriers): less waiting, more instruction streams.
  // mode 0: two contiguous halves; 1: warps of equal parity; 2: warps of equal (index mod 4), i.e. the warps of one scheduler
  static BT_DEV void group_sync(int mode) {
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5;
    int id, cnt;
    if (mode == 0) { const int h = nw >> 1; if (h == 0) return; id = w < h ? 1 : 2; cnt = w < h ? h : nw - h; }
    else { const int g = mode == 1 ? 2 : 4, r = w % g; id = 1 + r; cnt = (nw - r + g - 1) / g; }
    // immediate barrier ids (a register id would reserve all 16 hardware barriers)
    if (id == 1) asm volatile("bar.sync 1, %0;" ::"r"(cnt << 5) : "memory");
    else if (id == 2) asm volatile("bar.sync 2, %0;" ::"r"(cnt << 5) : "memory");
    else if (id == 3) asm volatile("bar.sync 3, %0;" ::"r"(cnt << 5) : "memory");
    else asm volatile("bar.sync 4, %0;" ::"r"(cnt << 5) : "memory");
  }
  // "any" over the warps of equal parity (group_sync mode 1) on the same named barriers
  static BT_DEV int group_any(int p) {
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, r = w & 1, cnt = ((nw - r + 1) >> 1) << 5;
    int out;
    if (r == 0) asm volatile("{\n.reg .pred q, o;\nsetp.ne.s32 q, %1, 0;\nbarrier.red.or.pred o, 1, %2, q;\nselp.s32 %0, 1, 0, o;\n}" : "=r"(out) : "r"(p), "r"(cnt) : "memory");
    else asm volatile("{\n.reg .pred q, o;\nsetp.ne.s32 q, %1, 0;\nbarrier.red.or.pred o, 2, %2, q;\nselp.s32 %0, 1, 0, o;\n}" : "=r"(out) : "r"(p), "r"(cnt) : "memory");
    return out;
  }
  // 8-lane groups: lane r (< 7) of a group holds u[0]; every lane of the group receives all seven values
  // (called by all 32 lanes: the chain loops of aba_factor are warp-uniform)
  template <int NR>
  static BT_DEV void gather7(const float* u, float* U, int lane) {
#pragma unroll
    for (int c = 0; c < 7; c++) U[c] = __shfl_sync(0xffffffffu, u[0], c, 8);  // width 8: lane c of the caller's own group
  }
};
#endif
template <>
struct BtLanes<1> {
  static BT_DEV void sync() {}
  static BT_DEV float allsum(float v) { return v; }
  template <int N>
  static BT_DEV void allsumN(float (&)[N], int) {}
  static BT_DEV int any(int p) { return p; }
  static BT_DEV int allmax(int v) { return v; }
  static BT_DEV void cta_sync() {}
  static BT_DEV void group_sync(int) {}
  static BT_DEV int cta_any(int p) { return p; }
  static BT_DEV int group_any(int p) { return p; }
  template <int NR>
  static BT_DEV void gather7(const float* u, float* U, int) {
    for (int c = 0; c < 7; c++) U[c] = u[c];
  }
};

// debug stop points for bt_forward_debug (parity tests of intermediates)
enum { BT_STOP_NONE = 0, BT_STOP_TREE = 1, BT_STOP_SMOOTH = 2, BT_STOP_M = 3, BT_STOP_FACTOR = 4, BT_STOP_QACC_SMOOTH = 5,
       BT_STOP_COLLISION = 6, BT_STOP_SOLVE = 7 };

template <int G, int DS, int CS>
struct BtEnv {
  typedef BtLanes<G> W;
  const BtDev& m;
  float* s;
  int lane;
  bool live;       // false: this lane group has no environment in this round (it only takes part in the CTA barriers)
  int niter;       // solver iterations of the last substep (diagnostic)
  float cdist[CS]; // contact distances of the last collision pass (diagnostic / tests)

  int obs_pad;     // 0..3, see obsbuf()
  int clip;        // this environment's reference clip (RodentMultiClip: rows clip * clip_len .. of the stacked clip tables)

  BT_DEV BtEnv(const BtDev& m_, float* s_, int lane_, bool live_ = true) : m(m_), s(s_), lane(lane_), live(live_), niter(0), obs_pad(0), clip(0) {}

  // ------------------------------------------------------------------ constant records shared by the warps of the CTA
  // N4 consecutive 128-bit words of the record table at offset `off` of sh_tab, starting at float `idx`: from shared memory when
  // the kernel staged the table (model.py: sh_stage_floats; tables are staged whole and in order, so "starts below the limit"
  // means staged), else through the read-only path.  Two explicit paths: a generic load could be neither hoisted above the
  // scratch stores nor served by the read-only path (measured: -2.5 % on the fly).
  // kSmallModel: the 2-slot variant serves the fly models (nv <= 64), whose tables fit the L1 (16 x 12 KB of scratch) and which
  // lost 1.5 - 2.4 % to the mere presence of the second path (same-box A/B, profiles/r2l_staged_records.txt): compiled out there,
  // and model.py stages nothing for them.
  static constexpr bool kSmallModel = DS <= 2;
  template <int N4>
  BT_DEV void crec(const float* tab, int off, int idx, float* o) const {   // tab = the table on its own (== sh_tab + off)
#ifdef __CUDACC__
    extern __shared__ __align__(16) float bt_cta_smem[];
    if (!kSmallModel && off < m.sh_stage_floats) {
#pragma unroll
      for (int q = 0; q < N4; q++) bt_ld4(bt_cta_smem + off + idx + 4 * q, o + 4 * q);
      return;
    }
#endif
#pragma unroll
    for (int q = 0; q < N4; q++) bt_ldg4(tab + idx + 4 * q, o + 4 * q);
  }
  BT_DEV void crec2(const float* tab, int off, int idx, float* o) const {   // two floats (8-byte aligned)
#ifdef __CUDACC__
    extern __shared__ __align__(16) float bt_cta_smem[];
    if (!kSmallModel && off < m.sh_stage_floats) {
      const float2 v = *reinterpret_cast<const float2*>(bt_cta_smem + off + idx);
      o[0] = v.x; o[1] = v.y;
      return;
    }
#endif
    bt_ldg2(tab + idx, o);
  }
  // ------------------------------------------------------------------ scratch regions
  BT_DEV float* qpos() const { return s + m.o_qpos; }
  BT_DEV float* qvel() const { return s + m.o_qvel; }
  BT_DEV float* act() const { return s + m.o_act; }
  BT_DEV float* ctrl() const { return s + m.o_ctrl; }
  BT_DEV float* warm() const { return s + m.o_warm; }
  BT_DEV float* xpos() const { return s + m.o_xpos; }
  BT_DEV float* xquat() const { return s + m.o_xquat; }
  // one 12-float record per dof: S_k = cdof_k (6) then G_k = U_k / D_k (6); 16-byte aligned so the chain sweeps load it
  // with three 128-bit shared-memory loads
  BT_DEV float* cdof() const { return s + m.o_cdof; }
  BT_DEV float* crb() const { return s + m.o_crb; }
  BT_DEV float* Dinv() const { return s + m.o_Dinv; }
  BT_DEV float* Dd() const { return s + m.o_Dd; }  // the pivots D_k themselves (mul_M through the factor)
  // ancestor-chain sums of cdof * v per contact body, by-products of the root->leaves sweeps: slot 0 v = qvel (velocity
  // sweep), 1 v = qacc_warmstart (mulM_down), 2 v = qacc_smooth (solve_down)
  BT_DEV float* cbJ(int slot) const { return s + m.o_cbJ + slot * 6 * m.ncb; }
  BT_DEV float* pvec() const { return s + m.o_pvec; }  // 6 per dof: sweep state of solve() / mul_M(); with the vectors behind it: cvel/cacc in the tree pass
  BT_DEV float* T() const { return s + m.o_T; }
  BT_DEV float* ref() const { return s + m.o_ref; }
  BT_DEV float* aforce() const { return s + m.o_aforce; }
  BT_DEV float* actdot() const { return s + m.o_actdot; }
  BT_DEV float* qfrc_smooth() const { return s + m.o_qfrc_smooth; }
  BT_DEV float* qacc_smooth() const { return s + m.o_qacc_smooth; }
  BT_DEV float* qacc() const { return s + m.o_qacc; }
  BT_DEV float* xv() const { return s + m.o_x; }
  BT_DEV float* search() const { return s + m.o_search; }
  BT_DEV float* qfrc_c() const { return s + m.o_qfrc_c; }
  // chain descriptor (model.py: chain_desc) and the 8-float slot that hands a chain's sweep state to its parent / children
  struct ChainD { int k0, kb, pc, c, cadr, nch, nseg, sadr, pdof; };
  BT_DEV ChainD desc_at(const int* d) const {
#ifdef __CUDACC__
    const int4 a = __ldg(reinterpret_cast<const int4*>(d));
    const int4 b = __ldg(reinterpret_cast<const int4*>(d + 4));
    return ChainD{a.x, a.y, a.z, a.w, b.x, b.y & 0xffff, b.y >> 16, b.z, b.w};
#else
    return ChainD{d[0], d[1], d[2], d[3], d[4], d[5] & 0xffff, d[5] >> 16, d[6], d[7]};
#endif
  }
  // descriptor of the chain that virtual lane `vl` (0..31) walks in pass `ps` of the one-lane-per-chain sweeps (kb < k0: none)
  BT_DEV ChainD pass_d(int ps, int vl) const { return desc_at(m.hpass_desc + 8 * (32 * ps + vl)); }
  BT_DEV float* ctop(int c) const { return pvec() + 8 * c; }
  // T-region views during the constraint phase
  BT_DEV float* congeo() const { return T(); }                          // [ncon][12] off(3) frame(9)
  BT_DEV float* wrench() const { return s + m.o_wrench; }               // [ncon + ncross][6]; over Dd | cbJ when they fit (model.py)
  BT_DEV float* cbA() const { return s + m.o_cbA; }                     // [ncb][6]

  // ================================================================== P1: forward tree pass
  // (MJX: smooth.kinematics, com_pos, com_vel, rne forward half, passive fluid), in five parts so that
  // only a frame composition (pose) and a 12-float recursion (velocity) remain serial:
  //  * body_frame  (lane per body)      the body's frame RELATIVE TO ITS PARENT after its own joints (sincos, local
  //                                      quaternion products), local joint anchors / axes;
  //  * compose     (lane per body chain) world pose = parent pose o local frame, carried in registers along the chain;
  //  * joint_cdof  (lane per joint)      world anchor / axis -> cdof (the S half of the dof records);
  //  * vel_sweep   (lane per dof chain)  cvel / cdof_dot / cacc, carried in registers along the chain;
  //  * body_local  (lane per body)       inertia about the reference point, RNE body force, fluid forces.
  BT_DEV void body_frame(int b) {
    // packed constants (model.py: body_rec = pos quat jntadr jntnum parent ref; jnt_rec = type qposadr dofadr at_origin pos
    // axis qpos0): three 128-bit loads each
    float br[12];
    crec<3>(m.body_rec, m.sho_body_rec, 12 * b, br);
    float p[3] = {br[0], br[1], br[2]};
    float q[4] = {br[3], br[4], br[5], br[6]};
    const int jadr = (int)br[7], jnum = (int)br[8];
    for (int jj = 0; jj < jnum; jj++) {
      float jr[12];
      crec<3>(m.jnt_rec, m.sho_jnt_rec, 12 * (jadr + jj), jr);
      const int qa = (int)jr[1], da = (int)jr[2];
      if ((int)jr[0] == BT_JNT_FREE) {
        // a free joint hangs off the world: the "local" frame is the absolute pose
        p[0] = qpos()[qa]; p[1] = qpos()[qa + 1]; p[2] = qpos()[qa + 2];
        q[0] = qpos()[qa + 3]; q[1] = qpos()[qa + 4]; q[2] = qpos()[qa + 5]; q[3] = qpos()[qa + 6];
        const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);  // x / norm(x), as mjx math.normalize
        const float inrm = bt_div(1.0f, nrm);
        q[0] *= inrm; q[1] *= inrm; q[2] *= inrm; q[3] *= inrm;
        // MJX kinematics stores the normalised quaternion back into qpos
        qpos()[qa + 3] = q[0]; qpos()[qa + 4] = q[1]; qpos()[qa + 5] = q[2]; qpos()[qa + 6] = q[3];
      } else {
        const float jp[3] = {jr[4], jr[5], jr[6]};
        const float ja[3] = {jr[7], jr[8], jr[9]};
        const bool at_origin = jr[3] != 0.f;  // model-uniform branch
        float anchor[3] = {p[0], p[1], p[2]}, axis[3], r[3], ql[4], q2[4];
        if (!at_origin) {
          bt_rotate(jp, q, r);
          anchor[0] += r[0]; anchor[1] += r[1]; anchor[2] += r[2];
        }
        bt_rotate(ja, q, axis);
        float* rec = cdof() + 12 * da;  // local axis / anchor, turned into cdof by joint_cdof()
        {
          const float r6[6] = {axis[0], axis[1], axis[2], anchor[0], anchor[1], anchor[2]};
          bt_st6(rec, r6);  // (scalar stores at the 12-float record stride are 4-way bank conflicts)
        }
        const float ang = 0.5f * (qpos()[qa] - jr[10]);
        float sn, cs_;
#ifdef __CUDACC__
        sincosf(ang, &sn, &cs_);
#else
        sn = sinf(ang); cs_ = cosf(ang);
#endif
        ql[0] = cs_; ql[1] = ja[0] * sn; ql[2] = ja[1] * sn; ql[3] = ja[2] * sn;
        bt_quat_mul(q, ql, q2);
        q[0] = q2[0]; q[1] = q2[1]; q[2] = q2[2]; q[3] = q2[3];
        if (!at_origin) {
          bt_rotate(jp, q, r);
          p[0] = anchor[0] - r[0]; p[1] = anchor[1] - r[1]; p[2] = anchor[2] - r[2];
        }
      }
    }
    if ((int)br[9] == 0) {
      // directly under the world: already a world pose; it is the reference point of its tree
      bt_quat_normalize(q);
      const int rs = (int)br[10];
      ref()[3 * rs] = p[0]; ref()[3 * rs + 1] = p[1]; ref()[3 * rs + 2] = p[2];
    }
    float* bp = pose_pos(m.nbanc & 1);
    float* bq = pose_quat(m.nbanc & 1);
#pragma unroll
    for (int k = 0; k < 3; k++) bp[3 * b + k] = p[k];
    bt_st4(bq + 4 * b, q);
  }
  // the two pose buffers of the pointer-jumping composition: 0 = (xpos, xquat), 1 = the T region (idle during the tree pass)
  BT_DEV float* pose_pos(int which) const { return which ? T() : xpos(); }
  BT_DEV float* pose_quat(int which) const { return which ? T() + ((3 * m.nbody + 3) & ~3) : xquat(); }  // 16-byte aligned

  BT_DEV void joint_cdof(int j) {
    const int b = BT_LDG(m.jnt_bodyid + j), p = BT_LDG(m.body_parentid + b), da = BT_LDG(m.jnt_dofadr + j);
    const int rs = BT_LDG(m.body_ref + b);
    const float rp[3] = {ref()[3 * rs], ref()[3 * rs + 1], ref()[3 * rs + 2]};
    if (BT_LDG(m.jnt_type + j) == BT_JNT_FREE) {
      float R[9];
      bt_quat_to_mat(xquat() + 4 * b, R);
      const float off[3] = {rp[0] - xpos()[3 * b], rp[1] - xpos()[3 * b + 1], rp[2] - xpos()[3 * b + 2]};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float c6[6] = {0.f, 0.f, 0.f, k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f};
        bt_st6(cdof() + 12 * (da + k), c6);
        float ax[3] = {R[k], R[3 + k], R[6 + k]}, lin[3];
        bt_cross(ax, off, lin);
        const float r6[6] = {ax[0], ax[1], ax[2], lin[0], lin[1], lin[2]};
        bt_st6(cdof() + 12 * (da + 3 + k), r6);
      }
      return;
    }
    float* rec = cdof() + 12 * da;
    float l6[6];
    bt_ld6(rec, l6);
    float al[3] = {l6[0], l6[1], l6[2]}, nl[3] = {l6[3], l6[4], l6[5]}, ax[3], an[3], lin[3];
    if (p == 0) {
      ax[0] = al[0]; ax[1] = al[1]; ax[2] = al[2]; an[0] = nl[0]; an[1] = nl[1]; an[2] = nl[2];
    } else {
      bt_rotate(al, xquat() + 4 * p, ax);
      bt_rotate(nl, xquat() + 4 * p, an);
      an[0] += xpos()[3 * p]; an[1] += xpos()[3 * p + 1]; an[2] += xpos()[3 * p + 2];
    }
    const float off[3] = {rp[0] - an[0], rp[1] - an[1], rp[2] - an[2]};
    bt_cross(ax, off, lin);
    const float r6[6] = {ax[0], ax[1], ax[2], lin[0], lin[1], lin[2]};
    bt_st6(rec, r6);
  }

  BT_DEV void body_local(int b) {
    const float* cv = pvec();
    float pos[3], quat[4], cvel[6], cacc[6], rp[3];
    float br[16];  // packed constants (model.py: bl_rec = ipos iquat inertia mass | fluidbox lastdof ref)
    crec<4>(m.bl_rec, m.sho_bl_rec, 16 * b, br);
    const int rs = (int)br[15];
#pragma unroll
    for (int k = 0; k < 3; k++) { pos[k] = xpos()[3 * b + k]; rp[k] = ref()[3 * rs + k]; }
    bt_ld4(xquat() + 4 * b, quat);
    const int ld = (int)br[14];  // last dof on the chain root -> body: the body moves with it
    if (ld >= 0) {
      float r12[12];
      bt_ld12(cv + 12 * ld, r12);
#pragma unroll
      for (int k = 0; k < 6; k++) { cvel[k] = r12[k]; cacc[k] = r12[6 + k]; }
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++) cvel[k] = 0.f;
      cacc[0] = cacc[1] = cacc[2] = 0.f;
      cacc[3] = -m.grav_x; cacc[4] = -m.grav_y; cacc[5] = -m.grav_z;
    }
    // body inertia about the tree reference point, world axes
    const float mass = br[10];
    float ci[10], cf[6];
    {
      float ip[3] = {br[0], br[1], br[2]};
      float iq[4] = {br[3], br[4], br[5], br[6]};
      float in[3] = {br[7], br[8], br[9]};
      float r[3], q2[4], R[9], off[3];
      bt_rotate(ip, quat, r);
      off[0] = pos[0] + r[0] - rp[0]; off[1] = pos[1] + r[1] - rp[1]; off[2] = pos[2] + r[2] - rp[2];
      bt_quat_mul(quat, iq, q2);
      bt_quat_to_mat(q2, R);
      const float oo = bt_dot3(off, off);
      ci[0] = R[0] * R[0] * in[0] + R[1] * R[1] * in[1] + R[2] * R[2] * in[2] + mass * (oo - off[0] * off[0]);
      ci[1] = R[3] * R[3] * in[0] + R[4] * R[4] * in[1] + R[5] * R[5] * in[2] + mass * (oo - off[1] * off[1]);
      ci[2] = R[6] * R[6] * in[0] + R[7] * R[7] * in[1] + R[8] * R[8] * in[2] + mass * (oo - off[2] * off[2]);
      ci[3] = R[0] * R[3] * in[0] + R[1] * R[4] * in[1] + R[2] * R[5] * in[2] - mass * off[0] * off[1];
      ci[4] = R[0] * R[6] * in[0] + R[1] * R[7] * in[1] + R[2] * R[8] * in[2] - mass * off[0] * off[2];
      ci[5] = R[3] * R[6] * in[0] + R[4] * R[7] * in[1] + R[5] * R[8] * in[2] - mass * off[1] * off[2];
      ci[6] = mass * off[0]; ci[7] = mass * off[1]; ci[8] = mass * off[2]; ci[9] = mass;
      float t1[6], t2[6], t3[6];
      bt_inert_mul(ci, cacc, t1);
      bt_inert_mul(ci, cvel, t2);
      bt_motion_cross_force(cvel, t2, t3);
#pragma unroll
      for (int k = 0; k < 6; k++) cf[k] = t1[k] + t3[k];
      // passive fluid forces (inertia-box model): a wrench on the body, folded into cfrc with opposite sign
      if ((m.density > 0.f || m.viscosity > 0.f) && mass > 0.f) {
        float box[3] = {br[11], br[12], br[13]};
        float c[3], lin[3], lang[3], llin[3], lf[6] = {0, 0, 0, 0, 0, 0};
        bt_cross(off, cvel, c);
        lin[0] = cvel[3] - c[0]; lin[1] = cvel[4] - c[1]; lin[2] = cvel[5] - c[2];
        bt_matT_vec(R, cvel, lang);
        bt_matT_vec(R, lin, llin);
        if (m.viscosity > 0.f) {
          const float diam = (box[0] + box[1] + box[2]) * (1.0f / 3.0f);
          const float ka = 3.14159265358979f * diam * diam * diam * m.viscosity, kl = 3.0f * 3.14159265358979f * diam * m.viscosity;
#pragma unroll
          for (int k = 0; k < 3; k++) { lf[k] = -lang[k] * ka; lf[3 + k] = -llin[k] * kl; }
        }
        if (m.density > 0.f) {
          const float b0 = box[0], b1 = box[1], b2 = box[2];
          const float q0 = b0 * b0 * b0 * b0, q1 = b1 * b1 * b1 * b1, q2_ = b2 * b2 * b2 * b2;
          const float sv[3] = {b1 * b2, b0 * b2, b0 * b1};
          const float sa[3] = {b0 * (q1 + q2_), b1 * (q0 + q2_), b2 * (q0 + q1)};
#pragma unroll
          for (int k = 0; k < 3; k++) {
            lf[3 + k] -= 0.5f * m.density * sv[k] * fabsf(llin[k]) * llin[k];
            lf[k] -= m.density * sa[k] * fabsf(lang[k]) * lang[k] * (1.0f / 64.0f);
          }
        }
        float tq[3], fc[3], oxf[3];
        bt_mat_vec(R, lf, tq);
        bt_mat_vec(R, lf + 3, fc);
        bt_cross(off, fc, oxf);
        cf[0] -= tq[0] + oxf[0]; cf[1] -= tq[1] + oxf[1]; cf[2] -= tq[2] + oxf[2];
        cf[3] -= fc[0]; cf[4] -= fc[1]; cf[5] -= fc[2];
      }
    }
    float* cr = crb() + 10 * b;
#pragma unroll
    for (int k = 0; k < 10; k++) cr[k] = ci[k];
    float* cfs = T() + 6 * b;
#pragma unroll
    for (int k = 0; k < 6; k++) cfs[k] = cf[k];
  }

  BT_DEV void tree_forward() {
    if (lane == 0) {
      xpos()[0] = xpos()[1] = xpos()[2] = 0.f;
      xquat()[0] = 1.f; xquat()[1] = xquat()[2] = xquat()[3] = 0.f;
    }
    for (int b = 1 + lane; b < m.nbody; b += G) body_frame(b);
    W::sync();
    // ---- compose: world pose = parent pose o local frame, by pointer jumping over the body tree: after round r every body
    // holds its pose relative to its 2^(r+1)-th ancestor (or the world), so ceil(log2(depth)) lane-parallel rounds replace a
    // 39-body serial chain.  Rounds ping-pong between the two pose buffers; the last one writes (xpos, xquat).
    for (int r = 0; r < m.nbanc; r++) {
      const float* sp = pose_pos((m.nbanc - r) & 1);
      const float* sq = pose_quat((m.nbanc - r) & 1);
      float* dp = pose_pos((m.nbanc - 1 - r) & 1);
      float* dq = pose_quat((m.nbanc - 1 - r) & 1);
      // work list of the round (model.py: cmp_item): compose with the 2^r-th ancestor, or copy a finished pose forward once
      for (int it = BT_LDG(m.cmp_adr + r) + lane, i1 = BT_LDG(m.cmp_adr + r + 1); it < i1; it += G) {
        const int w = BT_LDG(m.cmp_item + it);
        const int b = w & 0xfff, a = (w >> 12) & 0xfff;
        float pos[3] = {sp[3 * b], sp[3 * b + 1], sp[3 * b + 2]};
        float quat[4];
        bt_ld4(sq + 4 * b, quat);
        if (!(w & (1 << 29))) {
          const float pa[3] = {sp[3 * a], sp[3 * a + 1], sp[3 * a + 2]};
          float qa[4];
          bt_ld4(sq + 4 * a, qa);
          float rr[3], q2[4];
          bt_rotate(pos, qa, rr);
          pos[0] = pa[0] + rr[0]; pos[1] = pa[1] + rr[1]; pos[2] = pa[2] + rr[2];
          bt_quat_mul(qa, quat, q2);
          quat[0] = q2[0]; quat[1] = q2[1]; quat[2] = q2[2]; quat[3] = q2[3];
          if (w & (1 << 28)) bt_quat_normalize(quat);  // world pose reached
        }
#pragma unroll
        for (int k = 0; k < 3; k++) dp[3 * b + k] = pos[k];
        bt_st4(dq + 4 * b, quat);
      }
      W::sync();
    }
    for (int j = lane; j < m.njnt; j += G) joint_cdof(j);
    W::sync();
    // ---- velocity sweep on the dof chains: inclusive cvel / cacc per dof (12 floats) in the pvec.. region
    float* cv = pvec();
    for (int ps = 0; ps < m.nhpass; ps++) {
      for (int vl = lane; vl < 32; vl += G) {
        const ChainD chd = pass_d(ps, vl);
        if (chd.kb < chd.k0) continue;
        const int k0 = chd.k0, kb = chd.kb, par = chd.pdof;
        float cvel[6], cacc[6], snap[6];
        if (par >= 0) {
          float r12[12];
          bt_ld12(cv + 12 * par, r12);
#pragma unroll
          for (int i = 0; i < 6; i++) { cvel[i] = r12[i]; cacc[i] = r12[6 + i]; }
        } else {
#pragma unroll
          for (int i = 0; i < 6; i++) cvel[i] = 0.f;
          cacc[0] = cacc[1] = cacc[2] = 0.f;
          cacc[3] = -m.grav_x; cacc[4] = -m.grav_y; cacc[5] = -m.grav_z;
        }
        if (BT_LDG(m.dof_vflag + k0) == 0) {
          // hinge-only chain (every chain but a free root): tight loop, next record in flight, packed fp32 updates,
          // 128-bit stores of the 12-float (cvel | cacc) record
          float S[6], Sn[6], qv, qn = 0.f;
          bt_ld6(cdof() + 12 * k0, S);
          qv = qvel()[k0];
          for (int k = k0; k <= kb; k++) {
            if (k < kb) { bt_ld6(cdof() + 12 * (k + 1), Sn); qn = qvel()[k + 1]; }
            float cd[6];
            bt_motion_cross(cvel, S, cd);
            bt_axpy6(cacc, cd, qv);
            bt_axpy6(cvel, S, qv);
            bt_st12(cv + 12 * k, cvel, cacc);
#pragma unroll
            for (int i = 0; i < 6; i++) S[i] = Sn[i];
            qv = qn;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 6; i++) snap[i] = cvel[i];
          for (int k = k0; k <= kb; k++) {
            float S[6], cd[6];
            bt_ld6(cdof() + 12 * k, S);
            const float qv = qvel()[k];
            const int vf = BT_LDG(m.dof_vflag + k);  // 0 hinge, 1 free translation, 2 first / 3 later free rotation dof
            if (vf == 2) {
#pragma unroll
              for (int i = 0; i < 6; i++) snap[i] = cvel[i];
            }
            if (vf != 1) {
              // MuJoCo: the three rotational dofs of a free joint all use the velocity after its translational dofs
              bt_motion_cross(vf >= 2 ? snap : cvel, S, cd);
              bt_axpy6(cacc, cd, qv);
            }
            bt_axpy6(cvel, S, qv);
            bt_st12(cv + 12 * k, cvel, cacc);
          }
        }
      }
      W::sync();
    }
    // J qvel chain sums: cvel at the last dof of every contact body
    for (int it = lane; it < 6 * m.ncb; it += G) {
      const int cb = it / 6, j = it - 6 * cb;
      cbJ(0)[it] = cv[12 * BT_LDG(m.cb_lastdof + cb) + j];
    }
    for (int b = 1 + lane; b < m.nbody; b += G) body_local(b);
    W::sync();
    // link records: inertia (10) and RNE force (6) of all the bodies a dof carries, summed into the slot of its first body
    // (dof_irec) so that the serial sweeps read one record per dof
    if (m.nmerge > 0) {
      for (int it = lane; it < m.nmerge * 16; it += G) {
        const int r = it >> 4, j = it & 15;
        const int dst = BT_LDG(m.merge_dst + r);
        float* p = j < 10 ? crb() + 10 * dst + j : T() + 6 * dst + (j - 10);
        float acc = *p;
        for (int e = BT_LDG(m.merge_adr + r); e < BT_LDG(m.merge_adr + r + 1); e++) {
          const int src = BT_LDG(m.merge_src + e);
          acc += j < 10 ? crb()[10 * src + j] : T()[6 * src + (j - 10)];
        }
        *p = acc;
      }
      W::sync();
    }
  }

  // ================================================================== P3: actuation + smooth generalized forces
  BT_DEV void smooth_forces() {
    // transmission + fwd_actuation, one lane per actuator; the actuator's constants come as one packed
    // 16-float record (model.py: act_rec; absent limits are +-3e38, non-affine gain / bias terms are zeros)
    for (int u = lane; u < m.nu; u += G) {
      float r[16];
#pragma unroll
      for (int q = 0; q < 4; q++) bt_ldg4(m.act_rec + 16 * u + 4 * q, r + 4 * q);
      const float gear = r[0];
      float len = 0.f, vel = 0.f;
      for (int w = (int)r[14], w1 = w + (int)r[15]; w < w1; w++) {
        float wr[4];
        crec<1>(m.wrap_rec, m.sho_wrap_rec, 4 * w, wr);
        len += wr[0] * qpos()[(int)wr[1]];
        vel += wr[0] * qvel()[(int)wr[2]];
      }
      len *= gear; vel *= gear;
      const float c = bt_clampf(ctrl()[u], r[1], r[2]);
      float ca = c;
      const int aa = (int)r[11];
      if (aa >= 0) {
        ca = act()[aa];
        actdot()[aa] = (c - ca) / r[3];
      }
      const float gain = r[4] + (r[5] * len + r[6] * vel);
      const float bias = r[8] + r[9] * len + r[10] * vel;
      aforce()[u] = bt_clampf(gain * ca + bias, r[12], r[13]);
    }
    // tau = passive + actuator forces; the bias forces (RNE backward half) are subtracted inside aba_factor<true>, which
    // completes qfrc_smooth in place
    W::sync();
    for (int i = lane; i < m.nv; i += G) {
      float r[8];
      bt_ldg4(m.dof_rec + 8 * i, r);
      bt_ldg2(m.dof_rec + 8 * i + 4, r + 4);
      float f = -r[0] * qvel()[i];
      const int qa = (int)r[3];
      if (qa >= 0) f -= r[1] * (qpos()[qa] - r[2]);
      float fa = 0.f;
      for (int k = (int)r[4], k1 = k + (int)r[5]; k < k1; k++) {
        float ar[2];
        crec2(m.dofact_rec, m.sho_dofact_rec, 2 * k, ar);
        fa += ar[0] * aforce()[(int)ar[1]];
      }
      qfrc_smooth()[i] = f + fa;
    }
    W::sync();
  }

  // ================================================================== articulated-body factorisation (replaces qM / L'DL)
  // Eliminating the dofs of the tree-sparse qM from the leaves (MuJoCo's L'DL order) is the articulated-body recursion:
  //   A_k = I_k + sum_children (A_c - U_c U_c' / D_c),   U_k = A_k S_k,   D_k = S_k . U_k + armature_k (+ h * damping_k)
  // with S_k = cdof_k and I_k the spatial inertia of the bodies carried by dof k (6x6, all about the tree reference point, so
  // no frame transforms).  D_k are the pivots of L'DL and U_k . S_j its unscaled rows; neither qM nor the factor is ever
  // materialised.
  // Scheduling: the dof tree is cut into CHAINS (maximal single-child paths; consecutive dof ids by DFS numbering).  A chain
  // is walked sequentially with its state in registers and no synchronisation; only the chain tree (3 levels for the
  // rodent, 2 for the fly) needs warp syncs.  The 6x6 recursion uses one 8-lane group per chain (lane r owns row r).
  static constexpr int kGrp = G >= 8 ? 8 : 1;     // lanes per chain in aba_factor
  static constexpr int kNR = G >= 8 ? 1 : 8;      // rows per lane

  // Lane r < 6 of a chain's 8-lane group owns row r of the 6x6 articulated inertia A.  Row r of a link's spatial inertia
  // [[Ibar, [h]x], [-[h]x, m 1]] is six signed picks out of its 10 numbers: the per-row index / sign tables below turn
  // the 6x6 expansion into six lane-indexed shared-memory loads (no selects).
  // The two spare lanes of the group ride along on the same instruction stream (same loads, dot, downdate shapes):
  //   row 6: f_k, the RNE backward recursion (sum of the subtree's body forces; bias_k = S_k . f_k)      [kRne only]
  //   row 7: p_k, the leaves->root half of  X <- M^-1 X  for this very factor (u_k = X_k - S_k . p_k; p += G_k u_k),
  //          X = tau - bias (tau = passive + actuator forces, in qfrc_smooth, completed in place) when kRne, else xv.
  // so the bias forces and the first half of the solve cost no sweep of their own.  g_k = u_k / D_k goes to xv.
  template <bool kRne>
  BT_DEV void aba_factor(float hdamp) {
    const int grp = lane / kGrp, rl = lane % kGrp;
    constexpr int kNG = G / kGrp;  // chains in flight
    float* Ab = pvec();  // 48 floats per chain: rows 0..7 of the chain top, handed to the parent chain
    float* X = kRne ? qfrc_smooth() : xv();
    // pivot seeds armature_k + h * damping_k, replaced in place by 1 / D_k as the sweep passes
    for (int i = lane; i < m.nv; i += G) Dinv()[i] = BT_LDG(m.dof_armature + i) + hdamp * BT_LDG(m.dof_damping + i);
    // idx (4 bits each) and sign (2 bits each: 0 -> 0, 1 -> +1, 2 -> -1) of row r, packed
    const unsigned kIdx[6] = {0x780430u, 0x608513u, 0x067254u, 0x009780u, 0x090608u, 0x900067u};
    const unsigned kSgn[6] = {0x615u, 0x855u, 0x195u, 0x064u, 0x112u, 0x409u};
    int ix[kNR][6], rbase[kNR], rstride[kNR], pmul[kNR];
    float sg[kNR][6], rc[kNR][4];
#pragma unroll
    for (int i = 0; i < kNR; i++) {
      const int row = rl + i;
      const int r = row < 6 ? row : 0;
      // where the row's per-dof result goes (s[rbase + k * rstride]) and how it is formed: see the loop below
      rbase[i] = row < 6 ? m.o_cdof + 6 + row : (row == 6 ? (kRne ? m.o_qfrc_smooth : m.o_tmpv) : m.o_x);
      rstride[i] = row < 6 ? 12 : 1;
      pmul[i] = row == 6 ? 24 : 40;  // bytes per body-force / link-inertia record
      rc[i][0] = row == 7 ? 1.f : 0.f; rc[i][1] = row == 6 ? 0.f : 1.f;
      rc[i][2] = row < 6 ? 1.f : (row == 6 ? 0.f : -1.f); rc[i][3] = row == 6 ? 1.f : 0.f;
#ifdef __CUDACC__
      asm volatile("" : "+r"(pmul[i]));
      asm volatile("" : "+r"(rbase[i]), "+r"(rstride[i]), "+f"(rc[i][0]), "+f"(rc[i][1]), "+f"(rc[i][2]), "+f"(rc[i][3]));
#endif
      unsigned pi = kIdx[0], ps = kSgn[0];
#pragma unroll
      for (int q = 1; q < 6; q++) { pi = r == q ? kIdx[q] : pi; ps = r == q ? kSgn[q] : ps; }
#pragma unroll
      for (int j = 0; j < 6; j++) {
        ix[i][j] = 4 * ((pi >> (4 * j)) & 15);  // byte offset into the 10-float record
        const unsigned c = (ps >> (2 * j)) & 3;
        sg[i][j] = c == 0 ? 0.f : (c == 1 ? 1.f : -1.f);
        if (row == 6) { ix[i][j] = 4 * j; sg[i][j] = kRne ? 1.f : 0.f; }  // the 6-float body-force record
        if (row == 7) { ix[i][j] = 0; sg[i][j] = 0.f; }
        ix[i][j] += 4 * (row == 6 ? m.o_T : m.o_crb);  // byte offset from the scratch base; + rb * pmul selects the record
#ifdef __CUDACC__
        // ... made an ABSOLUTE 32-bit shared-memory address: the per-step address is then one multiply-add (rb * pmul + ix);
        // left relative to `s`, every step re-derived the shared window base (S2UR SR_CgaCtaId, ULEA, LEA) and added it six times
        ix[i][j] += (int)__cvta_generic_to_shared(s);
        // opaque to the optimiser: otherwise the decode above is rematerialised inside the per-dof loop
        asm volatile("" : "+r"(ix[i][j]), "+f"(sg[i][j]));
#endif
      }
    }
    W::sync();
    // The chain loops are warp-uniform (every lane runs the longest chain of the pass; shorter / absent chains are
    // predicated off), so the row exchange is a full-mask shuffle and there is no divergence bookkeeping.
    for (int ps = 0; ps < m.napass; ps++) {
      for (int vg = grp; vg < 4; vg += kNG) {  // one iteration on the device (four 8-lane groups); four on the 1-lane host build
        const ChainD cd = desc_at(m.apass_desc + 8 * (4 * ps + vg));
        const int c = cd.c, k0 = cd.k0, kb = cd.kb, cadr = cd.cadr, nch = cd.nch;
        const int maxlen = W::allmax(kb - k0 + 1);
        float a[kNR][6];
#pragma unroll
        for (int i = 0; i < kNR; i++)
#pragma unroll
          for (int j = 0; j < 6; j++) a[i][j] = 0.f;
        for (int e = 0; e < nch; e++) {
          {
            const float* cr = Ab + 48 * BT_LDG(m.cchild_id + cadr + e);
#pragma unroll
            for (int i = 0; i < kNR; i++)
#pragma unroll
              for (int j = 0; j < 6; j++) a[i][j] += cr[6 * (rl + i) + j];
          }
        }
        // running cursors of this group's current dof (they freeze on the last dof once the chain is exhausted, so
        // predicated-off lanes keep reading valid memory and store nothing)
        const int nstep = kb - k0 + 1;            // <= 0: no chain in this pass
        int kc = nstep > 0 ? kb : 0;
        const float* recp = cdof() + 12 * kc;
        const float* xp = X + kc;
        float* dp = Dinv() + kc;
        float* ddp = Dd() + kc;   // (its own cursor: addressed off `s`, the store re-derived the shared window base in every step)
        float* outp[kNR];
#pragma unroll
        for (int i = 0; i < kNR; i++) outp[i] = s + rbase[i] + kc * rstride[i];
        int rbn = nstep > 0 ? BT_LDG(m.dof_irec + kc) : -1;  // link record of the current dof, fetched one step ahead
        for (int t = 0; t < maxlen; t++) {
          const bool act = t < nstep;
          float S[6], u[kNR], U[7];
          const int rb = rbn;
          rbn = t + 1 < nstep ? BT_LDG(m.dof_irec + kc - 1) : -1;
          if (rb >= 0) {
#pragma unroll
            for (int i = 0; i < kNR; i++) {
              int off = rb * pmul[i];  // ix holds the region base too
#ifdef __CUDACC__
              asm volatile("" : "+r"(off));  // keep the product out of the six address computations (one 3-input add each)
#endif
#pragma unroll
              for (int j = 0; j < 6; j++)
#ifdef __CUDACC__
                a[i][j] += sg[i][j] * *reinterpret_cast<const float*>(__cvta_shared_to_generic((size_t)(unsigned)(off + ix[i][j])));
#else
                a[i][j] += sg[i][j] * *reinterpret_cast<const float*>(reinterpret_cast<const char*>(s) + off + ix[i][j]);
#endif
            }
          }
          bt_ld6(recp, S);
          const float Xk = *xp;
#pragma unroll
          for (int i = 0; i < kNR; i++) u[i] = bt_dot6(a[i], S);
          W::template gather7<kNR>(u, U, lane);  // U[0..5] = A S, U[6] = S . f = bias_k
          // lanes without a dof in this step get D = 1 (their result is discarded; 0 would only cost a denormal path)
          const float D = act ? *dp + bt_dot6(S, U) : 1.0f;
          const float inv = bt_rcp_pos(D);
          const float xk = kRne ? Xk - U[6] : Xk;
          if (act) {
#pragma unroll
            for (int i = 0; i < kNR; i++) {
              // rows 0..5: a_r -= (u_r / D) U;  row 6: untouched;  row 7: p += ((x_k - S . p) / D) U.  Branch-free: the three
              // kinds of rows share one instruction stream through per-row constants (c7 = [row 7], n6 = [row != 6]).
              const float ui = (u[i] - rc[i][0] * xk) * (inv * rc[i][1]);
              // one store per row: G_k[row] = U_row / D (rows 0..5); qfrc_smooth_k = x_k (row 6; a scratch slot when !kRne);
              // xv_k = g_k = u_k / D_k (row 7, consumed by solve_down)
              *outp[i] = rc[i][2] * ui + rc[i][3] * xk;
              if (rl + i == 0) { *dp = inv; if (kRne) *ddp = D; }   // (the pivots themselves are only read by M v, after the qM factor)
              bt_axpy6(a[i], U, -ui);
            }
          }
          if (t + 1 < nstep) {
            kc--; recp -= 12; xp--; dp--; ddp--;
#pragma unroll
            for (int i = 0; i < kNR; i++) outp[i] -= rstride[i];
          }
        }
        if (kb >= 0) {
#pragma unroll
          for (int i = 0; i < kNR; i++)
#pragma unroll
            for (int j = 0; j < 6; j++) Ab[48 * c + 6 * (rl + i) + j] = a[i][j];
        }
      }
      W::sync();
    }
  }

  // x <- M^-1 x  (M = qM + diag(h * damping) of the last aba_factor): the articulated-body solve, two O(nv) sweeps
  //   leaves->root (sweep_up<false>): p_k = sum_children pbar_c;  u_k = x_k - S_k . p_k;  pbar_k = p_k + G_k u_k   (G_k = U_k / D_k)
  //   root->leaves (solve_down):      a = a_parent;  x_k = u_k / D_k - G_k . a;  a_k = a + S_k x_k
  // one lane per chain, p / a carried in registers along the chain.  Run backwards, the same recursions apply M:
  //   root->leaves (mulM_down):       w_k = D_k (v_k + G_k . a);  a_k = a + S_k v_k
  //   leaves->root (sweep_up<true>):  y_k = w_k + S_k . q_k;  qbar_k = q_k + G_k w_k
  // a_k is the spatial acceleration of the bodies behind dof k: stored at the last dof of every contact body it IS the
  // ancestor-chain sum the constraint Jacobian needs (cbout).  (The first half of the two solves that follow a
  // factorisation directly -- smooth and Euler -- runs inside aba_factor itself.)
  // The per-dof loops are software-pipelined by hand (two register sets, the next dof's record is in flight while the
  // current one is consumed): the recursion is one dependent chain per lane, so shared-memory latency is otherwise exposed.
  // kMul = false: the leaves->root half of the solve, in place (u_k = x_k - S_k . p; x_k <- g_k = u_k / D_k; p += G_k u_k);
  // kMul = true: the leaves->root half of y = M v (y_k = w_k + S_k . q; q += G_k w_k), w from mulM_down
  template <bool kMul>
  BT_DEV void sweep_up(float* x, float* y) {
    for (int ps = m.nhpass - 1; ps >= 0; ps--) {
      for (int vl = lane; vl < 32; vl += G) {
        const ChainD cd = pass_d(ps, vl);
        if (cd.kb < cd.k0) continue;
        const int c = cd.c, k0 = cd.k0, kb = cd.kb;
        float p[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int e = 0; e < cd.nch; e++) {
          float t6[6];
          bt_ld6(ctop(BT_LDG(m.cchild_id + cd.cadr + e)), t6);
#pragma unroll
          for (int j = 0; j < 6; j++) p[j] += t6[j];
        }
        // running cursors at dof k: the pair (k, k - 1) is addressed with immediate offsets
        const float* rp = cdof() + 12 * kb;
        float* xq = x + kb;
        const float* dq = Dinv() + kb;
        float* yq = kMul ? y + kb : nullptr;
        auto load = [&](int o, float (&R)[12], float& xk, float& dk) {  // dof k - o
          bt_ld12(rp - 12 * o, R);
          xk = xq[-o];
          if (!kMul) dk = dq[-o];
        };
        auto step = [&](const float (&R)[12], float xk, float dk, int o) {
          if (kMul) {
            yq[-o] = xk + bt_dot6(R, p);
            bt_axpy6(p, R + 6, xk);
          } else {
            const float u = xk - bt_dot6(R, p);
            xq[-o] = u * dk;  // in place: g_k = u_k / D_k, consumed by the root->leaves pass
            bt_axpy6(p, R + 6, u);
          }
        };
        float A[12], B[12], xa, da = 0.f, xb, db = 0.f;
        int k = kb;
        load(0, A, xa, da);
        for (; k > k0; k -= 2) {
          load(1, B, xb, db);
          step(A, xa, da, 0);
          if (k - 2 >= k0) load(2, A, xa, da);
          step(B, xb, db, 1);
          rp -= 24; xq -= 2; dq -= 2;
          if (kMul) yq -= 2;
        }
        if (k == k0) step(A, xa, da, 0);
        bt_st6(ctop(c), p);
      }
      W::sync();
    }
  }
  // root->leaves sweeps: kMul = false: solve_down (x_k = g_k - G_k . a);  kMul = true: mulM_down (w_k = D_k (v_k + G_k . a)).
  // A chain is walked in SEGMENTS that end at the last dof of a contact body (seg_* tables), where `a` is written to cbout.
  template <bool kMul>
  BT_DEV void sweep_down(const float* in, const float* dscale, float* out, float* cbout) {
    for (int ps = 0; ps < m.nhpass; ps++) {
      for (int vl = lane; vl < 32; vl += G) {
        const ChainD cd = pass_d(ps, vl);
        if (cd.kb < cd.k0) continue;
        const int c = cd.c, k0 = cd.k0;
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (cd.pc >= 0) bt_ld6(ctop(cd.pc), a);
        // running cursors at dof k: the pair (k, k + 1) is addressed with immediate offsets
        const float* rp = cdof() + 12 * k0;
        const float* iq = in + k0;
        float* oq = out + k0;
        const float* dq = kMul ? dscale + k0 : nullptr;
        auto load = [&](int o, float (&R)[12], float& ik, float& dk) {  // dof k + o
          bt_ld12(rp + 12 * o, R);
          ik = iq[o];
          if (kMul) dk = dq[o];
        };
        auto step = [&](const float (&R)[12], float ik, float dk, int o) {
          if (kMul) {
            oq[o] = dk * (ik + bt_dot6(R + 6, a));
            bt_axpy6(a, R, ik);
          } else {
            const float xk = ik - bt_dot6(R + 6, a);  // x_k = (u_k - U_k . a) / D_k = g_k - G_k . a
            oq[o] = xk;
            bt_axpy6(a, R, xk);
          }
        };
        float A[12], B[12], ia, da = 0.f, ib, db = 0.f;
        int k = k0;
        for (int sgi = cd.sadr; sgi < cd.sadr + cd.nseg; sgi++) {
          const int ke = BT_LDG(m.seg_end + sgi);
          load(0, A, ia, da);
          for (; k < ke; k += 2) {
            load(1, B, ib, db);
            step(A, ia, da, 0);
            if (k + 2 <= ke) load(2, A, ia, da);
            step(B, ib, db, 1);
            rp += 24; iq += 2; oq += 2;
            if (kMul) dq += 2;
          }
          if (k == ke) {
            step(A, ia, da, 0);
            k++; rp += 12; iq++; oq++;
            if (kMul) dq++;
          }
          const int cbi = cbout ? BT_LDG(m.seg_cb + sgi) : -1;
          if (cbi >= 0) {
#pragma unroll
            for (int j = 0; j < 6; j++) cbout[6 * cbi + j] = a[j];
          }
        }
        bt_st6(ctop(c), a);
      }
      W::sync();
    }
  }
  // The two root->leaves sweeps that follow the qM factorisation -- w = mulM_down(v) (first half of qM v) and x = solve_down(g)
  // (second half of qM^-1 qfrc_smooth) -- need nothing but the factor, so they walk the chains TOGETHER: one pass over the dof
  // records, two independent recurrences per lane (a1 for M v, a2 for the solve: twice the instruction-level parallelism of a
  // lone sweep), both sets of contact-body chain sums dropped at the segment ends.  Hand-over slots: ctop(c) and ctop(nchain + c).
  BT_DEV void sweep_down_mul_and_solve(const float* v, const float* dscale, float* w, float* cb1, float* x, float* cb2) {
    for (int ps = 0; ps < m.nhpass; ps++) {
      for (int vl = lane; vl < 32; vl += G) {
        const ChainD cd = pass_d(ps, vl);
        if (cd.kb < cd.k0) continue;
        const int c = cd.c, k0 = cd.k0;
        float a1[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a2[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (cd.pc >= 0) { bt_ld6(ctop(cd.pc), a1); bt_ld6(ctop(m.nchain + cd.pc), a2); }
        const float* rp = cdof() + 12 * k0;
        const float* vq = v + k0;
        const float* dq = dscale + k0;
        float* wq = w + k0;
        float* xq = x + k0;
        auto load = [&](int o, float (&R)[12], float& vk, float& dk, float& gk) {  // dof k + o
          bt_ld12(rp + 12 * o, R);
          vk = vq[o]; dk = dq[o]; gk = xq[o];
        };
        auto step = [&](const float (&R)[12], float vk, float dk, float gk, int o) {
          wq[o] = dk * (vk + bt_dot6(R + 6, a1));
          const float xk = gk - bt_dot6(R + 6, a2);
          xq[o] = xk;
          bt_axpy6(a1, R, vk);
          bt_axpy6(a2, R, xk);
        };
        float A[12], B[12], va, da, ga, vb, db, gb;
        int k = k0;
        for (int sgi = cd.sadr; sgi < cd.sadr + cd.nseg; sgi++) {
          const int ke = BT_LDG(m.seg_end + sgi);
          load(0, A, va, da, ga);
          for (; k < ke; k += 2) {
            load(1, B, vb, db, gb);
            step(A, va, da, ga, 0);
            if (k + 2 <= ke) load(2, A, va, da, ga);
            step(B, vb, db, gb, 1);
            rp += 24; vq += 2; dq += 2; wq += 2; xq += 2;
          }
          if (k == ke) {
            step(A, va, da, ga, 0);
            k++; rp += 12; vq++; dq++; wq++; xq++;
          }
          const int cbi = BT_LDG(m.seg_cb + sgi);
          if (cbi >= 0) {
#pragma unroll
            for (int j = 0; j < 6; j++) { cb1[6 * cbi + j] = a1[j]; cb2[6 * cbi + j] = a2[j]; }
          }
        }
        bt_st6(ctop(c), a1);
        bt_st6(ctop(m.nchain + c), a2);
      }
      W::sync();
    }
  }
  BT_DEV void solve_down(float* x, float* cbout) { sweep_down<false>(x, nullptr, x, cbout); }  // in place: g -> M^-1 x
  // y = M v through the factor: mulM_down then mulM_up
  BT_DEV void mulM_down(const float* v, float* w, float* cbout) { sweep_down<true>(v, Dd(), w, cbout); }
  BT_DEV void mulM_up(float* w, float* y) { sweep_up<true>(w, y); }
  BT_DEV void solve(float* x, float* cbout) {
    sweep_up<false>(x, nullptr);
    solve_down(x, cbout);
  }

  // ================================================================== P8: collision (static contact list)
  BT_DEV void geom_pose(int g, float* gp, float* gR) {
    const int b = BT_LDG(m.cgeom_bodyid + g);
    float lp[3] = {BT_LDG(m.cgeom_pos + 3 * g), BT_LDG(m.cgeom_pos + 3 * g + 1), BT_LDG(m.cgeom_pos + 3 * g + 2)};
    float lq[4] = {BT_LDG(m.cgeom_quat + 4 * g), BT_LDG(m.cgeom_quat + 4 * g + 1), BT_LDG(m.cgeom_quat + 4 * g + 2),
                   BT_LDG(m.cgeom_quat + 4 * g + 3)};
    float r[3], q2[4];
    bt_rotate(lp, xquat() + 4 * b, r);
    gp[0] = xpos()[3 * b] + r[0]; gp[1] = xpos()[3 * b + 1] + r[1]; gp[2] = xpos()[3 * b + 2] + r[2];
    bt_quat_mul(xquat() + 4 * b, lq, q2);
    bt_quat_to_mat(q2, gR);
  }
  static BT_DEV void make_frame(const float* n, float* fr) {
    float a[3] = {n[0], n[1], n[2]}, b[3] = {0.f, 0.f, 0.f};
    if (a[1] > -0.5f && a[1] < 0.5f) b[1] = 1.f; else b[2] = 1.f;
    const float ab = bt_dot3(a, b);
    b[0] -= a[0] * ab; b[1] -= a[1] * ab; b[2] -= a[2] * ab;
    const float bn = 1.0f / sqrtf(bt_dot3(b, b));
    b[0] *= bn; b[1] *= bn; b[2] *= bn;
    fr[0] = a[0]; fr[1] = a[1]; fr[2] = a[2];
    fr[3] = b[0]; fr[4] = b[1]; fr[5] = b[2];
    bt_cross(a, b, fr + 6);
  }
  static BT_DEV void closest_seg(const float* a0, const float* a1, const float* b0, const float* b1, float* pa, float* pb) {
    float da[3], db[3], am[3], bm[3], tr[3];
    for (int k = 0; k < 3; k++) { da[k] = a1[k] - a0[k]; db[k] = b1[k] - b0[k]; }
    const float la = sqrtf(bt_dot3(da, da)), lb = sqrtf(bt_dot3(db, db));
    const float ia = la > 0.f ? bt_div(1.0f, la) : 1.0f, ib = lb > 0.f ? bt_div(1.0f, lb) : 1.0f;
    for (int k = 0; k < 3; k++) { da[k] *= ia; db[k] *= ib; }
    const float ha = 0.5f * la, hb = 0.5f * lb;
    for (int k = 0; k < 3; k++) { am[k] = a0[k] + da[k] * ha; bm[k] = b0[k] + db[k] * hb; tr[k] = am[k] - bm[k]; }
    const float dab = bt_dot3(da, db), dat = bt_dot3(da, tr), dbt = bt_dot3(db, tr);
    const float den = 1.f - dab * dab;
    const float ota = bt_div(-dat + dab * dbt, den + 1e-6f), otb = dbt + ota * dab;
    const float ta = bt_clampf(ota, -ha, ha), tb = bt_clampf(otb, -hb, hb);
    float ba[3], bb[3], na[3], nb[3], v[3];
    for (int k = 0; k < 3; k++) { ba[k] = am[k] + da[k] * ta; bb[k] = bm[k] + db[k] * tb; }
    for (int k = 0; k < 3; k++) v[k] = bb[k] - am[k];
    float t = bt_clampf(bt_dot3(v, da), -ha, ha);
    for (int k = 0; k < 3; k++) na[k] = am[k] + da[k] * t;
    for (int k = 0; k < 3; k++) v[k] = ba[k] - bm[k];
    t = bt_clampf(bt_dot3(v, db), -hb, hb);
    for (int k = 0; k < 3; k++) nb[k] = bm[k] + db[k] * t;
    float d1[3], d2[3];
    for (int k = 0; k < 3; k++) { d1[k] = na[k] - bb[k]; d2[k] = ba[k] - nb[k]; }
    if (bt_dot3(d1, d1) < bt_dot3(d2, d2)) { for (int k = 0; k < 3; k++) { pa[k] = na[k]; pb[k] = bb[k]; } }
    else { for (int k = 0; k < 3; k++) { pa[k] = ba[k]; pb[k] = nb[k]; } }
  }

  // fills congeo (contact point relative to the tree reference point + frame) and returns dist per lane-owned contact
  BT_DEV void collide() {
#pragma unroll
    for (int sl = 0; sl < CS; sl++) {
      const int c = lane + sl * G;
      cdist[sl] = 1.0f;
      if (c >= m.ncon) continue;
      const int g1 = BT_LDG(m.con_g1 + c), g2 = BT_LDG(m.con_g2 + c), fn = BT_LDG(m.con_fn + c);
      float p1[3], R1[9], p2[3], R2[9], pos[3], fr[9], dist;
      geom_pose(g1, p1, R1);
      geom_pose(g2, p2, R2);
      const float s20 = BT_LDG(m.cgeom_size + 3 * g2), s21 = BT_LDG(m.cgeom_size + 3 * g2 + 1), s22 = BT_LDG(m.cgeom_size + 3 * g2 + 2);
      if (fn == BT_FN_PLANE_CAPSULE) {
        const float n[3] = {R1[2], R1[5], R1[8]}, ax[3] = {R2[2], R2[5], R2[8]};
        const float na = bt_dot3(n, ax);
        float b[3] = {ax[0] - n[0] * na, ax[1] - n[1] * na, ax[2] - n[2] * na};
        const float bn = sqrtf(bt_dot3(b, b));
        if (bn < 0.5f) {
          const bool usey = n[1] > -0.5f && n[1] < 0.5f;
          b[0] = usey ? R1[1] : R1[2]; b[1] = usey ? R1[4] : R1[5]; b[2] = usey ? R1[7] : R1[8];
        } else {
          const float ib = bt_div(1.0f, bn);
          b[0] *= ib; b[1] *= ib; b[2] *= ib;
        }
        fr[0] = n[0]; fr[1] = n[1]; fr[2] = n[2]; fr[3] = b[0]; fr[4] = b[1]; fr[5] = b[2];
        bt_cross(n, b, fr + 6);
        const float sg = BT_LDG(m.con_sub + c) == 0 ? 1.f : -1.f;
        float sp[3] = {p2[0] + sg * ax[0] * s21, p2[1] + sg * ax[1] * s21, p2[2] + sg * ax[2] * s21};
        float df[3] = {sp[0] - p1[0], sp[1] - p1[1], sp[2] - p1[2]};
        dist = bt_dot3(df, n) - s20;
        for (int k = 0; k < 3; k++) pos[k] = sp[k] - n[k] * (s20 + 0.5f * dist);
      } else if (fn == BT_FN_PLANE_ELLIPSOID || fn == BT_FN_PLANE_SPHERE) {
        const float n[3] = {R1[2], R1[5], R1[8]};
        if (fn == BT_FN_PLANE_SPHERE) {
          float df[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
          dist = bt_dot3(df, n) - s20;
          for (int k = 0; k < 3; k++) pos[k] = p2[k] - n[k] * (s20 + 0.5f * dist);
        } else {
          float ln[3], sup[3], w[3];
          bt_matT_vec(R2, n, ln);
          sup[0] = ln[0] * s20; sup[1] = ln[1] * s21; sup[2] = ln[2] * s22;
          const float isn = 1.0f / sqrtf(bt_dot3(sup, sup));
          sup[0] = -sup[0] * isn * s20; sup[1] = -sup[1] * isn * s21; sup[2] = -sup[2] * isn * s22;
          bt_mat_vec(R2, sup, w);
          for (int k = 0; k < 3; k++) pos[k] = p2[k] + w[k];
          float df[3] = {pos[0] - p1[0], pos[1] - p1[1], pos[2] - p1[2]};
          dist = bt_dot3(df, n);
          for (int k = 0; k < 3; k++) pos[k] -= n[k] * dist * 0.5f;
        }
        make_frame(n, fr);
      } else {  // capsule-capsule
        const float s10 = BT_LDG(m.cgeom_size + 3 * g1), s11 = BT_LDG(m.cgeom_size + 3 * g1 + 1);
        float a0[3], a1[3], b0[3], b1[3], pa[3], pb[3];
        for (int k = 0; k < 3; k++) {
          a0[k] = p1[k] - R1[3 * k + 2] * s11; a1[k] = p1[k] + R1[3 * k + 2] * s11;
          b0[k] = p2[k] - R2[3 * k + 2] * s21; b1[k] = p2[k] + R2[3 * k + 2] * s21;
        }
        closest_seg(a0, a1, b0, b1, pa, pb);
        float n[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
        const float len = sqrtf(bt_dot3(n, n));
        if (len < BT_MINVAL) { n[0] = 1.f; n[1] = 0.f; n[2] = 0.f; }
        else { const float il = bt_div(1.0f, len); n[0] *= il; n[1] *= il; n[2] *= il; }
        dist = len - (s10 + s20);
        for (int k = 0; k < 3; k++) pos[k] = pa[k] + n[k] * (s10 + 0.5f * dist);
        make_frame(n, fr);
      }
      cdist[sl] = dist;
      const int rs = BT_LDG(m.con_ref + c);
      const float off3[3] = {pos[0] - ref()[3 * rs], pos[1] - ref()[3 * rs + 1], pos[2] - ref()[3 * rs + 2]};
      const float r6a[6] = {off3[0], off3[1], off3[2], fr[0], fr[1], fr[2]}, r6b[6] = {fr[3], fr[4], fr[5], fr[6], fr[7], fr[8]};
      bt_st12(congeo() + 12 * c, r6a, r6b);
    }
    W::sync();
  }

  // ================================================================== matrix-free constraint Jacobian
  // out[sl][k] = frame_k . (J_point(body2) - J_point(body1)) v   for lane-owned contacts, from the per-contact-body
  // ancestor-chain sums `cbs` = sum_{d in chain(cb)} cdof_d v_d left behind by the root->leaves sweep that handled v
  BT_DEV void jproj(const float* cbs, float out[CS][3]) {
#pragma unroll
    for (int sl = 0; sl < CS; sl++) {
      const int c = lane + sl * G;
      out[sl][0] = out[sl][1] = out[sl][2] = 0.f;
      if (c >= m.ncon) continue;
      const int cb1 = BT_LDG(m.con_cb1 + c), cb2 = BT_LDG(m.con_cb2 + c);
      float cg[12];  // contact record (offset 3, frame 9): three 128-bit loads (scalar loads at stride 12 are 4-way conflicts)
      bt_ld12(congeo() + 12 * c, cg);
      float w[3];
      const int xr = m.ncross > 0 ? BT_LDG(m.con_xref + c) : -1;
      if (xr >= 0) {
        // contact between two kinematic trees: each chain sum is about its own tree's reference point, so the point velocity
        // of body 1 uses the offset from ITS reference: off1 = off + ref[tree of body 2] - ref[tree of body 1]
        const int r2 = BT_LDG(m.con_ref + c);
        const float off1[3] = {cg[0] + (ref()[3 * r2] - ref()[3 * xr]), cg[1] + (ref()[3 * r2 + 1] - ref()[3 * xr + 1]),
                               cg[2] + (ref()[3 * r2 + 2] - ref()[3 * xr + 2])};
        float w1[3];
        bt_cross(cbs + 6 * cb2, cg, w);
        bt_cross(cbs + 6 * cb1, off1, w1);
#pragma unroll
        for (int k = 0; k < 3; k++) w[k] = (w[k] + cbs[6 * cb2 + 3 + k]) - (w1[k] + cbs[6 * cb1 + 3 + k]);
      } else {
        float A[6] = {0, 0, 0, 0, 0, 0};
        if (cb2 >= 0) for (int k = 0; k < 6; k++) A[k] += cbs[6 * cb2 + k];
        if (cb1 >= 0) for (int k = 0; k < 6; k++) A[k] -= cbs[6 * cb1 + k];
        bt_cross(A, cg, w);
        w[0] += A[3]; w[1] += A[4]; w[2] += A[5];
      }
      out[sl][0] = bt_dot3(cg + 3, w); out[sl][1] = bt_dot3(cg + 6, w); out[sl][2] = bt_dot3(cg + 9, w);
    }
  }

  // ================================================================== constraint rows (registers)
  struct Efc {
    float jar[CS][4];  // J qacc - aref per row
    float D[CS][4];    // 0 => row absent / contact inactive
    float mu[CS][2];
    float ljar[DS];    // joint-limit rows (one per lane-owned dof)
    float lD[DS];
    float lsg[DS];     // +1 / -1; 0 => no row
  };

  BT_DEV void zero_rows(Efc& e) const {
#pragma unroll
    for (int sl = 0; sl < CS; sl++) {
#pragma unroll
      for (int r = 0; r < 4; r++) { e.jar[sl][r] = 0.f; e.D[sl][r] = 0.f; }
      e.mu[sl][0] = e.mu[sl][1] = 0.f;
    }
#pragma unroll
    for (int sl = 0; sl < DS; sl++) { e.ljar[sl] = 0.f; e.lD[sl] = 0.f; e.lsg[sl] = 0.f; }
  }

  BT_DEV void kbi(float sr0, float sr1, const float* si, float pos, float& k, float& b, float& imp) const {
    float timeconst = sr0, dampratio = sr1;
    float dmin = BT_LDG(si), dmax = BT_LDG(si + 1), width = BT_LDG(si + 2), mid = BT_LDG(si + 3), power = BT_LDG(si + 4);
    if (timeconst < 2.f * m.timestep) timeconst = 2.f * m.timestep;  // refsafe
    dmin = bt_clampf(dmin, BT_MINIMP, BT_MAXIMP);
    dmax = bt_clampf(dmax, BT_MINIMP, BT_MAXIMP);
    if (width < BT_MINVAL) width = BT_MINVAL;
    mid = bt_clampf(mid, BT_MINIMP, BT_MAXIMP);
    if (power < 1.f) power = 1.f;
    k = bt_div(1.f, dmax * dmax * timeconst * timeconst * dampratio * dampratio);
    b = bt_div(2.f, dmax * timeconst);
    if (sr0 <= 0.f) k = bt_div(-sr0, dmax * dmax);
    if (sr1 <= 0.f) b = bt_div(-sr1, dmax);
    const float x = bt_div(fabsf(pos), width);
    float y;
    if (power == 2.f) {  // MuJoCo default; the general case goes through the out-of-line helper (code size)
      y = x < mid ? bt_div(x * x, mid) : 1.f - bt_div((1.f - x) * (1.f - x), 1.f - mid);
    } else {
      y = bt_impedance_pow(x, mid, power);
    }
    float im = dmin + y * (dmax - dmin);
    im = bt_clampf(im, dmin, dmax);
    if (x > 1.f) im = dmax;
    imp = im;
  }

  // make_constraint: fills D / mu / limit rows and the row-wise aref (returned for the jar initialisation)
  BT_DEV void make_rows(Efc& e, const float jv[CS][3], float aref[CS][4], float laref[DS]) {
#pragma unroll
    for (int sl = 0; sl < CS; sl++) {
      const int c = lane + sl * G;
#pragma unroll
      for (int r = 0; r < 4; r++) { e.D[sl][r] = 0.f; e.jar[sl][r] = 0.f; aref[sl][r] = 0.f; }
      e.mu[sl][0] = e.mu[sl][1] = 0.f;
      if (c >= m.ncon) continue;
      const float pos = cdist[sl] - BT_LDG(m.con_includemargin + c);
      if (!(pos < 0.f)) continue;
      float k, b, imp;
      kbi(BT_LDG(m.con_solref + 2 * c), BT_LDG(m.con_solref + 2 * c + 1), m.con_solimp + 5 * c, pos, k, b, imp);
      const float t = BT_LDG(m.con_invweight + c);
      const int dim = BT_LDG(m.con_dim + c);
      if (dim == 1) {
        float R = bt_div(t * (1.f - imp), imp);
        R = R < BT_MINVAL ? BT_MINVAL : R;
        e.D[sl][0] = bt_div(1.f, R);
        aref[sl][0] = -b * jv[sl][0] - k * imp * pos;
      } else if (m.cone == BT_CONE_PYRAMIDAL) {
#pragma unroll
        for (int a = 0; a < 2; a++) {
          const float mu = BT_LDG(m.con_mu + 2 * c + a);
          e.mu[sl][a] = mu;
          float R = bt_div(bt_div((t + mu * mu * t) * 2.f * mu * mu, m.impratio) * (1.f - imp), imp);
          R = R < BT_MINVAL ? BT_MINVAL : R;
          e.D[sl][2 * a] = e.D[sl][2 * a + 1] = bt_div(1.f, R);
          aref[sl][2 * a] = -b * (jv[sl][0] + mu * jv[sl][1 + a]) - k * imp * pos;
          aref[sl][2 * a + 1] = -b * (jv[sl][0] - mu * jv[sl][1 + a]) - k * imp * pos;
        }
      } else {
        // elliptic: rows 0 (normal), 1, 2 (friction); friction rows: pos 0 in aref, impedance from the normal
        e.mu[sl][0] = BT_LDG(m.con_mu + 2 * c); e.mu[sl][1] = BT_LDG(m.con_mu + 2 * c + 1);
        float R = bt_div(t * (1.f - imp), imp);
        R = R < BT_MINVAL ? BT_MINVAL : R;
        e.D[sl][0] = bt_div(1.f, R);
        float Rf = bt_div(bt_div(t, m.impratio) * (1.f - imp), imp);
        Rf = Rf < BT_MINVAL ? BT_MINVAL : Rf;
        e.D[sl][1] = e.D[sl][2] = bt_div(1.f, Rf);
        aref[sl][0] = -b * jv[sl][0] - k * imp * pos;
        aref[sl][1] = -b * jv[sl][1];
        aref[sl][2] = -b * jv[sl][2];
      }
    }
#pragma unroll
    for (int sl = 0; sl < DS; sl++) {
      const int i = lane + sl * G;
      e.lD[sl] = 0.f; e.lsg[sl] = 0.f; e.ljar[sl] = 0.f; laref[sl] = 0.f;
      if (i >= m.nv || !BT_LDG(m.dof_limited + i)) continue;
      const float q = qpos()[BT_LDG(m.dof_qposadr + i)];
      const float dlo = q - BT_LDG(m.dof_range + 2 * i), dhi = BT_LDG(m.dof_range + 2 * i + 1) - q;
      const float pos = (dlo < dhi ? dlo : dhi) - BT_LDG(m.dof_margin + i);
      if (!(pos < 0.f)) continue;
      const float sg = dlo < dhi ? 1.f : -1.f;
      float k, b, imp;
      kbi(BT_LDG(m.dof_solref + 2 * i), BT_LDG(m.dof_solref + 2 * i + 1), m.dof_solimp + 5 * i, pos, k, b, imp);
      float R = bt_div(BT_LDG(m.dof_invweight0 + i) * (1.f - imp), imp);
      R = R < BT_MINVAL ? BT_MINVAL : R;
      e.lD[sl] = bt_div(1.f, R);
      e.lsg[sl] = sg;
      laref[sl] = -b * sg * qvel()[i] - k * imp * pos;
    }
  }

  // base (normal, t1, t2) projections -> row values
  BT_DEV void row_combine(const Efc& e, int sl, const float base[3], float out[4]) const {
    if (m.cone == BT_CONE_PYRAMIDAL) {
      out[0] = base[0] + e.mu[sl][0] * base[1]; out[1] = base[0] - e.mu[sl][0] * base[1];
      out[2] = base[0] + e.mu[sl][1] * base[2]; out[3] = base[0] - e.mu[sl][1] * base[2];
    } else {
      out[0] = base[0]; out[1] = base[1]; out[2] = base[2]; out[3] = 0.f;
    }
  }

  // elliptic cone helper: zone classification shared by cost / force / line search (oracle update_constraint)
  // returns 0 top, 1 bottom, 2 middle
  BT_DEV int ell_zone(const Efc& e, int sl, const float ja[3], float& mu0, float& Dm, float& N, float& Tn, float u[3]) const {
    const float Dn = e.D[sl][0];
    mu0 = e.mu[sl][0] * sqrtf(Dn / e.D[sl][1]);
    Dm = Dn / (mu0 * mu0 * (1.f + mu0 * mu0));
    u[0] = ja[0] * mu0; u[1] = ja[1] * e.mu[sl][0]; u[2] = ja[2] * e.mu[sl][1];
    N = u[0];
    Tn = sqrtf(u[1] * u[1] + u[2] * u[2]);
    if (N >= mu0 * Tn || (Tn <= 0.f && N >= 0.f)) return 0;
    if (mu0 * N + Tn <= 0.f || (Tn <= 0.f && N < 0.f)) return 1;
    return 2;
  }

  // cost of the constraint rows at jar; optionally the base-direction forces per contact and limit forces
  template <bool want_force>
  BT_DEV float rows_cost(const Efc& e, float fbase[CS][3], float lforce[DS]) const {
    float cost = 0.f;
#pragma unroll
    for (int sl = 0; sl < CS; sl++) {
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      if (m.cone == BT_CONE_PYRAMIDAL || e.D[sl][1] == 0.f) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const float ja = e.jar[sl][r];
          if (ja < 0.f) { f[r] = -e.D[sl][r] * ja; cost += 0.5f * e.D[sl][r] * ja * ja; }
        }
        if (want_force) {
          fbase[sl][0] = f[0] + f[1] + f[2] + f[3];
          fbase[sl][1] = e.mu[sl][0] * (f[0] - f[1]);
          fbase[sl][2] = e.mu[sl][1] * (f[2] - f[3]);
          if (m.cone != BT_CONE_PYRAMIDAL) { fbase[sl][0] = f[0]; fbase[sl][1] = fbase[sl][2] = 0.f; }
        }
      } else {
        float mu0, Dm, N, Tn, u[3];
        const int z = ell_zone(e, sl, e.jar[sl], mu0, Dm, N, Tn, u);
        if (z == 1) {
#pragma unroll
          for (int r = 0; r < 3; r++) { const float ja = e.jar[sl][r]; f[r] = -e.D[sl][r] * ja; cost += 0.5f * e.D[sl][r] * ja * ja; }
        } else if (z == 2) {
          const float NmT = N - mu0 * Tn;
          cost += 0.5f * Dm * NmT * NmT;
          f[0] = -Dm * NmT * mu0;
          f[1] = Dm * NmT * mu0 / Tn * u[1] * e.mu[sl][0];
          f[2] = Dm * NmT * mu0 / Tn * u[2] * e.mu[sl][1];
        }
        if (want_force) { fbase[sl][0] = f[0]; fbase[sl][1] = f[1]; fbase[sl][2] = f[2]; }
      }
    }
#pragma unroll
    for (int sl = 0; sl < DS; sl++) {
      const float ja = e.ljar[sl];
      float f = 0.f;
      if (ja < 0.f) { f = -e.lD[sl] * ja; cost += 0.5f * e.lD[sl] * ja * ja; }
      if (want_force) lforce[sl] = f;
    }
    return W::allsum(cost);
  }

  // qfrc_constraint = J^T f  (per-dof gather over the contacts whose chain contains the dof) -> scratch + regs
  BT_DEV void jt_force(const Efc& e, const float fbase[CS][3], const float lforce[DS]) {
#ifdef __CUDACC__
    if (G == 32 && CS == 1 && m.jt_seg_steps > 0) {
      // Every contact has ONE moving body and the contacts of a contact body sit in consecutive lanes (model.py: jt_seg; the
      // rodent: 30 floor contacts on 8 bodies, segments of 1-6): the wrench stays in registers and the per-body sums are a
      // segmented shuffle reduction (3 steps) instead of a shared-memory round trip with dependent index loads.
      float w6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int mask = 0, head = -1;
      if (lane < m.ncon) {
        float cg[12];
        bt_ld12(congeo() + 12 * lane, cg);
        float F[3], tq[3];
#pragma unroll
        for (int k = 0; k < 3; k++) F[k] = cg[3 + k] * fbase[0][0] + cg[6 + k] * fbase[0][1] + cg[9 + k] * fbase[0][2];
        bt_cross(cg, F, tq);
        w6[0] = tq[0]; w6[1] = tq[1]; w6[2] = tq[2]; w6[3] = F[0]; w6[4] = F[1]; w6[5] = F[2];
        const int sm_ = BT_LDG(m.con_seg + lane);   // bits 0..4: lane + 2^s is in the same segment; bits 8..: contact body + 1 when this lane heads it
        mask = sm_ & 0xff;
        head = (sm_ >> 8) - 1;
      }
      for (int st = 0; st < m.jt_seg_steps; st++) {
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const float o = __shfl_down_sync(0xffffffffu, w6[j], 1 << st);
          if ((mask >> st) & 1) w6[j] += o;
        }
      }
      if (head >= 0) {
#pragma unroll
        for (int j = 0; j < 6; j++) cbA()[6 * head + j] = w6[j];
      }
      W::sync();
    } else
#endif
    {
#pragma unroll
    for (int sl = 0; sl < CS; sl++) {
      const int c = lane + sl * G;
      if (c >= m.ncon) continue;
      float cg[12];
      bt_ld12(congeo() + 12 * c, cg);
      float F[3], tq[3];
#pragma unroll
      for (int k = 0; k < 3; k++) F[k] = cg[3 + k] * fbase[sl][0] + cg[6 + k] * fbase[sl][1] + cg[9 + k] * fbase[sl][2];
      bt_cross(cg, F, tq);
      float* w = wrench() + 6 * c;
      w[0] = tq[0]; w[1] = tq[1]; w[2] = tq[2]; w[3] = F[0]; w[4] = F[1]; w[5] = F[2];
      if (m.ncross > 0) {
        const int xr = BT_LDG(m.con_xref + c);
        if (xr >= 0) {
          // the same force on body 1's tree, as a wrench about THAT tree's reference point, in the contact's second slot
          // (model.py: con_xslot; the contact-body sums of body 1 read it with a minus sign)
          const int r2 = BT_LDG(m.con_ref + c);
          const float off1[3] = {cg[0] + (ref()[3 * r2] - ref()[3 * xr]), cg[1] + (ref()[3 * r2 + 1] - ref()[3 * xr + 1]),
                                 cg[2] + (ref()[3 * r2 + 2] - ref()[3 * xr + 2])};
          float t1[3];
          bt_cross(off1, F, t1);
          float* w1 = wrench() + 6 * BT_LDG(m.con_xslot + c);
          w1[0] = t1[0]; w1[1] = t1[1]; w1[2] = t1[2]; w1[3] = F[0]; w1[4] = F[1]; w1[5] = F[2];
        }
      }
    }
    W::sync();
    // wrench per contact BODY (8 for the rodent instead of 30 contacts), then one gather per dof over the contact bodies
    // whose ancestor chain contains the dof
    for (int it = lane; it < m.ncb * 6; it += G) {
      const int cb = it / 6, j = it - cb * 6;
      float acc = 0.f;
      for (int k = BT_LDG(m.cbcon_adr + cb), k1 = BT_LDG(m.cbcon_adr + cb + 1); k < k1; k++) {
        const int cs = BT_LDG(m.cbcon_cs + k);  // contact index, sign in the top bit (body of geom1: -1)
        const float w = wrench()[6 * (cs & 0x7fffffff) + j];
        acc += cs < 0 ? -w : w;
      }
      cbA()[it] = acc;
    }
    W::sync();
    }
    // one summed wrench per GROUP of dofs with the same contact bodies below them (in the per-contact wrench slots, which
    // are dead by now), then one dot product per dof
    float* wg = wrench();
    for (int it = lane; it < m.nwgrp * 6; it += G) {
      const int g = it / 6, j = it - g * 6;
      float acc = 0.f;
      if (!kSmallModel && m.wgrp_contig) {
        // the contact bodies below a dof are a contiguous range (DFS numbering): one packed load, then independent shared loads
        const int rg = BT_LDG(m.wgrp_rng + g);
        const float* p = cbA() + 6 * (rg & 0xffff) + j;
        for (int k = rg >> 16; k > 0; k--, p += 6) acc += *p;
      } else {
        for (int k = BT_LDG(m.wgrp_adr + g), k1 = BT_LDG(m.wgrp_adr + g + 1); k < k1; k++) acc += cbA()[6 * BT_LDG(m.wgrp_cb + k) + j];
      }
      wg[it] = acc;
    }
    W::sync();
#pragma unroll
    for (int sl = 0; sl < DS; sl++) {
      const int i = lane + sl * G;
      if (i < m.nv) {
        float acc = e.lsg[sl] * lforce[sl];
        const int g = BT_LDG(m.dof_wgrp + i);
        if (g >= 0) {
          float S[6];
          bt_ld6(cdof() + 12 * i, S);
          acc += bt_dot6(S, wg + 6 * g);
        }
        qfrc_c()[i] = acc;
      }
    }
    W::sync();
  }

  // ------------------------------------------------------------------ exact 1-D line search (MJX solver._linesearch)
  struct LsPt { float alpha, cost, d0, d1; };

  // evaluates NP trial step sizes at once: per lane partial (q0,q1,q2) sums over its rows, then warp sums
  template <int NP>
  BT_DEV void ls_eval(const Efc& e, const float jv[CS][4], const float ljv[DS], const float qg[3], const float* alpha, LsPt* out) const {
    float q0[NP], q1[NP], q2[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) { q0[p] = q1[p] = q2[p] = 0.f; }
#pragma unroll
    for (int sl = 0; sl < CS; sl++) {
      if (m.cone == BT_CONE_PYRAMIDAL || e.D[sl][1] == 0.f) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const float D = e.D[sl][r], ja = e.jar[sl][r], v = jv[sl][r];
          const float a0 = 0.5f * D * ja * ja, a1 = D * v * ja, a2 = 0.5f * D * v * v;
#pragma unroll
          for (int p = 0; p < NP; p++)
            if (ja + alpha[p] * v < 0.f) { q0[p] += a0; q1[p] += a1; q2[p] += a2; }
        }
      } else if (e.D[sl][0] != 0.f) {
        const float Dn = e.D[sl][0];
        const float mu0 = e.mu[sl][0] * sqrtf(Dn / e.D[sl][1]);
        const float Dm = Dn / (mu0 * mu0 * (1.f + mu0 * mu0));
        const float u0 = e.jar[sl][0] * mu0, v0 = jv[sl][0] * mu0;
        const float u1 = e.jar[sl][1] * e.mu[sl][0], v1 = jv[sl][1] * e.mu[sl][0];
        const float u2 = e.jar[sl][2] * e.mu[sl][1], v2 = jv[sl][2] * e.mu[sl][1];
        const float uu = u1 * u1 + u2 * u2, uv = u1 * v1 + u2 * v2, vv = v1 * v1 + v2 * v2;
#pragma unroll
        for (int p = 0; p < NP; p++) {
          const float al = alpha[p];
          const float N = u0 + al * v0;
          const float Tsq = uu + al * (2.f * uv + al * vv);
          const float Tn = sqrtf(Tsq > 0.f ? Tsq : 0.f);
          if (N >= mu0 * Tn || (Tn <= 0.f && N >= 0.f)) {
          } else if (mu0 * N + Tn <= 0.f || (Tn <= 0.f && N < 0.f)) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
              const float D = e.D[sl][r], ja = e.jar[sl][r], v = jv[sl][r];
              q0[p] += 0.5f * D * ja * ja; q1[p] += D * v * ja; q2[p] += 0.5f * D * v * v;
            }
          } else {
            const float N1 = v0, T1 = (uv + al * vv) / Tn;
            const float T2 = vv / Tn - (uv + al * vv) * T1 / (Tn * Tn);
            const float NmT = N - mu0 * Tn, dd = N1 - mu0 * T1;
            const float c0 = 0.5f * Dm * NmT * NmT, c1 = Dm * NmT * dd, c2 = Dm * (dd * dd + NmT * (-mu0 * T2));
            q0[p] += c0 - c1 * al + 0.5f * c2 * al * al;
            q1[p] += c1 - c2 * al;
            q2[p] += 0.5f * c2;
          }
        }
      }
    }
#pragma unroll
    for (int sl = 0; sl < DS; sl++) {
      const float D = e.lD[sl], ja = e.ljar[sl], v = ljv[sl];
      const float a0 = 0.5f * D * ja * ja, a1 = D * v * ja, a2 = 0.5f * D * v * v;
#pragma unroll
      for (int p = 0; p < NP; p++)
        if (ja + alpha[p] * v < 0.f) { q0[p] += a0; q1[p] += a1; q2[p] += a2; }
    }
    float red[3 * NP];
#pragma unroll
    for (int p = 0; p < NP; p++) { red[3 * p] = q0[p]; red[3 * p + 1] = q1[p]; red[3 * p + 2] = q2[p]; }
    W::template allsumN<3 * NP>(red, lane);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      const float s0 = red[3 * p] + qg[0], s1 = red[3 * p + 1] + qg[1], s2 = red[3 * p + 2] + qg[2];
      const float al = alpha[p];
      out[p].alpha = al;
      out[p].cost = al * al * s2 + al * s1 + s0;
      out[p].d0 = 2.f * al * s2 + s1;
      out[p].d1 = 2.f * s2 + (s2 == 0.f ? BT_MINVAL : 0.f);
    }
  }

  // ================================================================== P9-P11: constraint solve (CG, primal in qacc)
  // in: qfrc_smooth, qacc_smooth, warm (scratch), Ma_warm = M * warm in `qfrc_c` (scratch), L^-1 in LD
  // out: qacc, qfrc_c (scratch); warm <- qacc
  // Same algorithm as MJX solver.solve; the loop is rotated so that update_constraint /
  // update_gradient / the M^-1 solve / the line search each have ONE call site (instruction-cache footprint).
  BT_DEV void solve_constraints() {
    Efc e;
    float qfs[DS], qas[DS], Ma[DS];
    zero_rows(e);
#pragma unroll
    for (int sl = 0; sl < DS; sl++) qfs[sl] = qas[sl] = Ma[sl] = 0.f;
    if (live) {
#pragma unroll
      for (int sl = 0; sl < DS; sl++) {
        const int i = lane + sl * G;
        qfs[sl] = i < m.nv ? qfrc_smooth()[i] : 0.f;
        qas[sl] = i < m.nv ? qacc_smooth()[i] : 0.f;
        Ma[sl] = i < m.nv ? qfrc_c()[i] : 0.f;  // M * warm for now
      }
    }
    float gauss = 0.f;
    if (live) {
      // candidate loop: c = 0 rows from qvel (make_constraint), c = 1 cost at qacc_warmstart, c = 2 cost at qacc_smooth
      float aref[CS][4], laref[DS], jw[CS][4], ljw[DS];
      float cost_w = 0.f, gauss_w = 0.f;
      for (int c = 0; c < 3; c++) {
        const float* vec = c == 0 ? qvel() : (c == 1 ? warm() : qacc_smooth());
        float jb[CS][3];
        jproj(cbJ(c), jb);
        if (c == 0) {
          make_rows(e, jb, aref, laref);
          continue;
        }
#pragma unroll
        for (int sl = 0; sl < CS; sl++) {
          row_combine(e, sl, jb[sl], e.jar[sl]);
#pragma unroll
          for (int r = 0; r < 4; r++) e.jar[sl][r] -= aref[sl][r];
        }
        float g = 0.f;
#pragma unroll
        for (int sl = 0; sl < DS; sl++) {
          const int i = lane + sl * G;
          const float vi = i < m.nv ? vec[i] : 0.f;
          e.ljar[sl] = e.lsg[sl] * vi - laref[sl];
          if (c == 1) g += (Ma[sl] - qfs[sl]) * (vi - qas[sl]);
        }
        const float rc = rows_cost<false>(e, nullptr, nullptr);
        if (c == 1) {
          gauss_w = 0.5f * W::allsum(g);
          cost_w = rc + gauss_w;
#pragma unroll
          for (int sl = 0; sl < CS; sl++)
#pragma unroll
            for (int r = 0; r < 4; r++) jw[sl][r] = e.jar[sl][r];
#pragma unroll
          for (int sl = 0; sl < DS; sl++) ljw[sl] = e.ljar[sl];
        } else {
          // warm-start selection (MJX solver.solve): keep the cheaper of qacc_warmstart / qacc_smooth
          const bool use_warm = cost_w < rc;
          if (use_warm) {
#pragma unroll
            for (int sl = 0; sl < CS; sl++)
#pragma unroll
              for (int r = 0; r < 4; r++) e.jar[sl][r] = jw[sl][r];
#pragma unroll
            for (int sl = 0; sl < DS; sl++) e.ljar[sl] = ljw[sl];
          }
#pragma unroll
          for (int sl = 0; sl < DS; sl++) {
            const int i = lane + sl * G;
            if (!use_warm) Ma[sl] = qfs[sl];  // M * qacc_smooth = qfrc_smooth
            if (i < m.nv) qacc()[i] = use_warm ? warm()[i] : qas[sl];
          }
          gauss = use_warm ? gauss_w : 0.f;
        }
      }
      W::sync();
    }
    float grad[DS], Mgrad[DS], mv[DS];
#pragma unroll
    for (int sl = 0; sl < DS; sl++) grad[sl] = Mgrad[sl] = mv[sl] = 0.f;
    float sb[CS][3];  // J * search in the contact frames, carried by the same recurrence as the search direction
#pragma unroll
    for (int sl = 0; sl < CS; sl++) sb[sl][0] = sb[sl][1] = sb[sl][2] = 0.f;
    float cost = INFINITY, prev_cost = 0.f;
    const float nvf = (float)(m.nv > 1 ? m.nv : 1);
    const float scale = 1.0f / (m.meaninertia * nvf);
    int it = 0;
    bool active = live;
    float pg_pMg = 0.f, g_pMg = 0.f;
    while (true) {
      if (active) {
      // ---- update_constraint: cost + forces + qfrc_constraint = J' f
      {
        float fbase[CS][3], lforce[DS];
        const float c = rows_cost<true>(e, fbase, lforce);
        jt_force(e, fbase, lforce);
        prev_cost = cost;
        cost = c + gauss;
      }
      // ---- update_gradient: grad = M a - qfrc_smooth - qfrc_constraint; Mgrad = M^-1 grad; Polak-Ribiere direction
      pg_pMg = 0.f; g_pMg = 0.f;
      float gnorm2 = 0.f;
#pragma unroll
      for (int sl = 0; sl < DS; sl++) {
        const int i = lane + sl * G;
        pg_pMg += grad[sl] * Mgrad[sl];
        grad[sl] = i < m.nv ? Ma[sl] - qfs[sl] - qfrc_c()[i] : 0.f;
        g_pMg += grad[sl] * Mgrad[sl];
        gnorm2 += grad[sl] * grad[sl];
        if (i < m.nv) xv()[i] = grad[sl];
      }
      gnorm2 = W::allsum(gnorm2);
      // ---- termination test (MJX solve.cond); it does not involve Mgrad, so the M^-1 solve of the final pass is skipped
      {
        const float improvement = (prev_cost - cost) * scale;
        const float gradient = sqrtf(gnorm2) * scale;
        if (it >= m.iterations || improvement < m.tolerance || gradient < m.tolerance) active = false;
      }
      }
      if (active) {
      W::sync();
      solve(xv(), cbA());  // leaves the ancestor-chain sums of Mgrad in cbA
      float g_Mg = 0.f;
#pragma unroll
      for (int sl = 0; sl < DS; sl++) {
        const int i = lane + sl * G;
        Mgrad[sl] = i < m.nv ? xv()[i] : 0.f;
        g_Mg += grad[sl] * Mgrad[sl];
      }
      { float r3[3] = {pg_pMg, g_pMg, g_Mg}; if (G > 1) { W::template allsumN<3>(r3, lane); pg_pMg = r3[0]; g_pMg = r3[1]; g_Mg = r3[2]; } }
      float beta = 0.f;
      if (it > 0) {
        beta = bt_div(g_Mg - g_pMg, pg_pMg > BT_MINVAL ? pg_pMg : BT_MINVAL);
        beta = beta < 0.f ? 0.f : beta;
      }
#pragma unroll
      for (int sl = 0; sl < DS; sl++) {
        const int i = lane + sl * G;
        mv[sl] = -grad[sl] + beta * mv[sl];  // M * search without a product: M * Mgrad = grad
        if (i < m.nv) search()[i] = -Mgrad[sl] + (it > 0 ? beta * search()[i] : 0.f);
      }
      W::sync();
      // ---- exact line search along `search` (MJX solver._linesearch)
      float jv[CS][4], ljv[DS];
      {
        // search = -Mgrad + beta * search  =>  J search = -J Mgrad + beta * J search
        float mb[CS][3];
        jproj(cbA(), mb);
#pragma unroll
        for (int sl = 0; sl < CS; sl++) {
#pragma unroll
          for (int k = 0; k < 3; k++) sb[sl][k] = -mb[sl][k] + (it > 0 ? beta * sb[sl][k] : 0.f);
          row_combine(e, sl, sb[sl], jv[sl]);
        }
      }
      float sn = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
      for (int sl = 0; sl < DS; sl++) {
        const int i = lane + sl * G;
        const float sv = i < m.nv ? search()[i] : 0.f;
        ljv[sl] = e.lsg[sl] * sv;
        sn += sv * sv;
        g1 += sv * (Ma[sl] - qfs[sl]);
        g2 += 0.5f * sv * mv[sl];
      }
      { float r3[3] = {sn, g1, g2}; if (G > 1) { W::template allsumN<3>(r3, lane); sn = r3[0]; g1 = r3[1]; g2 = r3[2]; } }
      const float qg[3] = {gauss, g1, g2};
      const float gtol = m.tolerance * m.ls_tolerance * sqrtf(sn) * m.meaninertia * nvf;
      LsPt p0, lo, hi;
      p0.alpha = p0.cost = p0.d0 = p0.d1 = 0.f;
      lo = hi = p0;
      // stage 0: p0 = point(0); stage 1: lo = point(newton step from p0); stage >= 2: bracketing iterations
      bool swap = true;
      // stages 0 / 1 evaluate a single point each (one call site of the 1-point evaluator)
      for (int stage = 0; stage < 2; stage++) {
        const float a1 = stage == 0 ? 0.f : p0.alpha - p0.d0 * bt_rcp_pos(p0.d1);
        LsPt pt1;
        ls_eval<1>(e, jv, ljv, qg, &a1, &pt1);
        if (stage == 0) p0 = pt1;
        else if (pt1.d0 < p0.d0) { lo = pt1; hi = p0; }
        else { lo = p0; hi = pt1; }
      }
      for (int stage = 2;; stage++) {
        float al[3];
        {
          bool done = stage - 2 >= m.ls_iterations;
          done |= !swap;
          done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
          done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
          if (done) break;
          // d1 = 2 s2 (+ MINVAL) > 0: Newton-refined reciprocal (<= 1 ulp) instead of the IEEE division sequence
          al[0] = lo.alpha - lo.d0 * bt_rcp_pos(lo.d1); al[1] = hi.alpha - hi.d0 * bt_rcp_pos(hi.d1); al[2] = 0.5f * (lo.alpha + hi.alpha);
        }
        LsPt pt[3];
        ls_eval<3>(e, jv, ljv, qg, al, pt);
        const bool s_lo_next = (lo.d0 > 0.f) || (lo.d0 < pt[0].d0);
        if (s_lo_next) lo = pt[0];
        const bool s_lo_mid = (pt[2].d0 < 0.f) && (lo.d0 < pt[2].d0);
        if (s_lo_mid) lo = pt[2];
        const bool s_hi_next = (hi.d0 < 0.f) || (hi.d0 > pt[1].d0);
        if (s_hi_next) hi = pt[1];
        const bool s_hi_mid = (pt[2].d0 > 0.f) && (hi.d0 > pt[2].d0);
        if (s_hi_mid) hi = pt[2];
        swap = s_lo_next || s_lo_mid || s_hi_next || s_hi_mid;
      }
      const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
      const float alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
      if (improved) {
#pragma unroll
        for (int sl = 0; sl < DS; sl++) {
          const int i = lane + sl * G;
          if (i < m.nv) qacc()[i] += search()[i] * alpha;
          Ma[sl] += mv[sl] * alpha;
          e.ljar[sl] += ljv[sl] * alpha;
        }
#pragma unroll
        for (int sl = 0; sl < CS; sl++)
#pragma unroll
          for (int r = 0; r < 4; r++) e.jar[sl][r] += jv[sl][r] * alpha;
      }
      W::sync();
      float g = 0.f;
#pragma unroll
      for (int sl = 0; sl < DS; sl++) {
        const int i = lane + sl * G;
        if (i < m.nv) g += (Ma[sl] - qfs[sl]) * (qacc()[i] - qas[sl]);
      }
      gauss = 0.5f * W::allsum(g);
      it++;
      }
      // the warps share this loop's code whatever their iteration; with sync bit 16 they also walk it in step (a warp that
      // has converged idles through the remaining passes: it would wait at the next substep barrier anyway)
      if (m.sync_mode & 16) { if (!((m.sync_mode & 1024) ? W::group_any(active) : W::cta_any(active))) break; }
      else if (!active) break;
    }
    niter = it;
    if (live) for (int i = lane; i < m.nv; i += G) warm()[i] = qacc()[i];
    W::sync();
  }

  // ================================================================== mjx.step = forward (phase 0) + euler (phase 1)
  // Both phases share ONE call site of aba_factor / solve: phase 0 uses qM, phase 1 qM + h * diag(damping) (MJX euler
  // with implicit joint damping).
  // returns false when stopped early by a debug stop point.
  BT_DEV bool substep(bool do_euler, int stop = BT_STOP_NONE, int frame = 0) {
    // `live` is warp-uniform and `stop` / `do_euler` are CTA-uniform, so every warp of the CTA reaches every barrier
    // sync_mode bits: 1 substep start, 2 after the tree pass, 4 before each factorisation, 8 before collision, 16 every CG pass
    //                 (1024: the extra points 2..16 align the warps of equal parity only, as bit 64 does at the substep start)
    const int sm = m.sync_mode;
    if (sm & 1) W::cta_sync();
    if (sm & 32) W::group_sync(0);
    if (sm & 64) W::group_sync(1);
    if (sm & 128) W::group_sync(2);
    if (live) tree_forward();
    if (stop == BT_STOP_TREE) return false;
    if (sm & 2) { if (sm & 1024) W::group_sync(1); else W::cta_sync(); }
    if (live) smooth_forces();
    if (stop == BT_STOP_SMOOTH) return false;
    const float h = m.timestep;
    for (int phase = 0; phase < (do_euler ? 2 : 1); phase++) {
      if ((sm & 4) && !(sm & (phase == 0 ? 4096 : 2048))) { if (sm & 1024) W::group_sync(1); else W::cta_sync(); }  // 2048 / 4096: first / second only
      if (live) {
        if (phase == 0) {
          // factor qM; the spare rows of the sweep finish qfrc_smooth (RNE bias) and run the first half of the smooth solve
          aba_factor<true>(0.f);
        } else {
          // MJX euler with implicit joint damping: (qM + h diag(damping)) qacc = qfrc_smooth + qfrc_constraint
          for (int i = lane; i < m.nv; i += G) xv()[i] = qfrc_smooth()[i] + qfrc_c()[i];
          aba_factor<false>(h);
        }
      }
      if (stop == BT_STOP_M || stop == BT_STOP_FACTOR) return false;
      if (live) {
        if (phase == 0) {
          // qM * qacc_warmstart (consumed by the solver's warm-start test) through the factor
          // ... together with the root->leaves half of the smooth solve (both need only the factor)
          sweep_down_mul_and_solve(warm(), Dd(), qacc(), cbJ(1), xv(), cbJ(2));
          mulM_up(qacc(), qfrc_c());
        } else {
          solve_down(xv(), nullptr);
        }
      }
      if (phase == 0) {
        if (live) {
          for (int i = lane; i < m.nv; i += G) qacc_smooth()[i] = xv()[i];
          W::sync();
        }
        if (stop == BT_STOP_QACC_SMOOTH) return false;
        if (sm & 8) { if (sm & 1024) W::group_sync(1); else W::cta_sync(); }
        if (live) collide();
        if (stop == BT_STOP_COLLISION) return false;
        solve_constraints();
      }
    }
    if (do_euler && live) integrate();
    return true;
  }
  BT_DEV bool forward(int stop = BT_STOP_NONE) { return substep(false, stop); }
  BT_DEV void step(int frame = 0) { substep(true, BT_STOP_NONE, frame); }

  // mjx _advance with qacc = xv (the implicitly damped acceleration)
  BT_DEV void integrate() {
    const float h = m.timestep;
    for (int u = lane; u < m.na; u += G) act()[u] += h * actdot()[u];
    for (int i = lane; i < m.nv; i += G) {
      const float v = qvel()[i] + h * xv()[i];
      qvel()[i] = v;
      const int qa = BT_LDG(m.dof_qposadr + i);
      if (qa >= 0) qpos()[qa] += h * v;
    }
    W::sync();
    for (int j = lane; j < m.njnt; j += G) {
      if (BT_LDG(m.jnt_type + j) != BT_JNT_FREE) continue;
      const int qa = BT_LDG(m.jnt_qposadr + j), da = BT_LDG(m.jnt_dofadr + j);
      float* q = qpos() + qa;
      const float* v = qvel() + da;
      q[0] += h * v[0]; q[1] += h * v[1]; q[2] += h * v[2];
      const float w[3] = {v[3], v[4], v[5]};
      const float n = sqrtf(bt_dot3(w, w));
      float ax[3] = {0.f, 0.f, 0.f};
      if (n > 0.f) { const float in = bt_div(1.0f, n); ax[0] = w[0] * in; ax[1] = w[1] * in; ax[2] = w[2] * in; }
      const float ha = 0.5f * h * n;
      const float sn = sinf(ha), cs = cosf(ha);
      float qr[4] = {cs, ax[0] * sn, ax[1] * sn, ax[2] * sn}, q2[4];
      bt_quat_mul(q + 3, qr, q2);
      bt_quat_normalize(q2);
      q[3] = q2[0]; q[4] = q2[1]; q[5] = q2[2]; q[6] = q2[3];
    }
    W::sync();
  }

  // initcheck substitute (compute-sanitizer is closed on the pool): with the table scalar `poison` set, every program starts
  // by filling its scratch slice with NaN, so a read of anything the program did not write itself shows up in the outputs
  // (tests/test_gpu_parity.py::test_poisoned_scratch_and_scheduling_invariance requires bit-identical results)
  BT_DEV void poison_scratch() {
    if (!m.poison) return;
    for (int i = lane; i < m.smem_floats; i += G) s[i] = NAN;
    W::sync();
  }

  // ================================================================== state I/O ([n_envs, dim] rows, coalesced per env)
  BT_DEV void load_state(const BtState& st, int env) {
    for (int i = lane; i < m.nq; i += G) qpos()[i] = st.qpos[(size_t)env * m.nq + i];
    for (int i = lane; i < m.nv; i += G) { qvel()[i] = st.qvel[(size_t)env * m.nv + i]; warm()[i] = st.qacc_warmstart[(size_t)env * m.nv + i]; }
    for (int i = lane; i < m.na; i += G) act()[i] = st.act[(size_t)env * m.na + i];
    W::sync();
  }
  BT_DEV void store_state(const BtState& st, int env, float time) {
    for (int i = lane; i < m.nq; i += G) st.qpos[(size_t)env * m.nq + i] = qpos()[i];
    for (int i = lane; i < m.nv; i += G) { st.qvel[(size_t)env * m.nv + i] = qvel()[i]; st.qacc_warmstart[(size_t)env * m.nv + i] = warm()[i]; }
    for (int i = lane; i < m.na; i += G) st.act[(size_t)env * m.na + i] = act()[i];
    if (st.xpos) for (int i = lane; i < 3 * m.nbody; i += G) st.xpos[(size_t)env * 3 * m.nbody + i] = xpos()[i];
    if (lane == 0) st.time[env] = time;
  }

  // ================================================================== env layer
  // staged observation row (crb/LD/T are dead by then), shifted by obs_pad floats so that it sits at the same offset modulo 16 bytes
  // as its destination row (bt_write_obs: 128-bit copies)
  BT_DEV float* obsbuf() const { return s + m.o_crb + obs_pad; }

  // per-animal constants of the env layer (model.py: animal_rec = qadr dadr nj jbase torso 0 0 0): one record per tracked
  // free root.  Every reference env has ONE animal; the two-rodent stress model (BASELINE.json configs[3]) has two, each
  // tracking its own copy of the clip (DESIGN.md "config 4").  The tethered fly is one animal with qadr = 0 and nj = nq.
  struct Animal { int qadr, dadr, nj, jbase, torso; };
  BT_DEV Animal animal(int a) const {
    const int* r = m.animal_rec + 8 * a;
    return Animal{BT_LDG(r), BT_LDG(r + 1), BT_LDG(r + 2), BT_LDG(r + 3), BT_LDG(r + 4)};
  }

  // _get_obs (fruitfly.py:598-646 free root; :271-319 tethered) into obsbuf, nan_to_num applied.
  // Layout: qpos | qvel | per animal: track_pos_local (3 L) | quat_dist (4 L) | joint_dist (n_joint_idxs L) | body_pos_dist_local (3 n_body_idxs L)
  BT_DEV void build_obs(int cur_frame) {
    float* o = obsbuf();
    const int L = m.ref_len, nq = m.nq, nv = m.nv, NA = m.n_animals;
    int start = cur_frame + 1;
    start = start < 0 ? 0 : (start > m.clip_len - L ? m.clip_len - L : start);  // dynamic_slice clamps the start
    start += clip * m.clip_len;                                                 // row of the stacked multi-clip tables
    for (int i = lane; i < nq; i += G) o[i] = bt_nan_to_num(qpos()[i]);
    for (int i = lane; i < nv; i += G) o[nq + i] = bt_nan_to_num(qvel()[i]);
    int base = nq + nv;
    const int nj = m.clip_nj;
    for (int a = 0; a < NA; a++) {
      const Animal an = animal(a);
      const float* qa = qpos() + an.qadr;
      const float rq[4] = {qa[3], qa[4], qa[5], qa[6]};  // tethered quirk: four joint angles (fruitfly.py:305)
      if (m.free_jnt) {
        for (int l = lane; l < L; l += G) {
          const float* cp = m.clip_position + 3 * ((start + l) * NA + a);
          float d[3] = {BT_LDG(cp) - qa[0], BT_LDG(cp + 1) - qa[1], BT_LDG(cp + 2) - qa[2]}, r[3];
          bt_rotate(d, rq, r);
          o[base + 3 * l] = bt_nan_to_num(r[0]); o[base + 3 * l + 1] = bt_nan_to_num(r[1]); o[base + 3 * l + 2] = bt_nan_to_num(r[2]);
          const float* cq = m.clip_quaternion + 4 * ((start + l) * NA + a);
          const float tq[4] = {BT_LDG(cq), BT_LDG(cq + 1), BT_LDG(cq + 2), BT_LDG(cq + 3)};
          const float inv[4] = {rq[0], -rq[1], -rq[2], -rq[3]};
          float rel[4];
          bt_quat_mul(tq, inv, rel);  // relative_quat(q_root, q_ref) = q_ref * conj(q_root)
#pragma unroll
          for (int k = 0; k < 4; k++) o[base + 3 * L + 4 * l + k] = bt_nan_to_num(rel[k]);
        }
        base += 7 * L;
      }
      const int qoff = an.qadr + (m.free_jnt ? 7 : 0);
      const int j0 = BT_LDG(m.jidx_adr + a), nji = BT_LDG(m.jidx_adr + a + 1) - j0;
      for (int it = lane; it < L * nji; it += G) {
        const int l = it / nji, k = it - l * nji;
        const int col = BT_LDG(m.joint_idxs + j0 + k);
        o[base + it] = bt_nan_to_num(BT_LDG(m.clip_joints + (size_t)(start + l) * nj + an.jbase + col) - qpos()[qoff + col]);
      }
      base += L * nji;
      const int b0 = BT_LDG(m.bidx_adr + a), nbi = BT_LDG(m.bidx_adr + a + 1) - b0;
      for (int it = lane; it < L * nbi; it += G) {
        const int l = it / nbi, k = it - l * nbi;
        const int b = BT_LDG(m.body_idxs + b0 + k);
        const float* cb = m.clip_body_positions + ((size_t)(start + l) * m.nbody + b) * 3;
        float d[3] = {BT_LDG(cb) - xpos()[3 * b], BT_LDG(cb + 1) - xpos()[3 * b + 1], BT_LDG(cb + 2) - xpos()[3 * b + 2]}, r[3];
        bt_rotate(d, rq, r);
        o[base + 3 * it] = bt_nan_to_num(r[0]); o[base + 3 * it + 1] = bt_nan_to_num(r[1]); o[base + 3 * it + 2] = bt_nan_to_num(r[2]);
      }
      base += 3 * L * nbi;
    }
    W::sync();
  }

  struct StepOut {
    float reward, done, metrics[BT_NMETRIC], summed_pos, quat_d, joint_d;
    int cur_frame, steps_taken;
  };

  // everything in env.step after pipeline_step (fruitfly.py:502-592); `action` = this env's row
  BT_DEV void reward_terms(const float* action, int cur_frame_in, int steps_taken_in, StepOut& r) {
    int stc = steps_taken_in + 1;
    const int hit = (stc == m.steps_for_cur_frame) ? 1 : 0;
    const int cur = cur_frame_in + hit;
    stc = stc * (hit ? 0 : 1);
    r.cur_frame = cur; r.steps_taken = stc;
    const int fi = (cur < 0 ? 0 : (cur > m.clip_len - 1 ? m.clip_len - 1 : cur)) + clip * m.clip_len;  // JAX gather clamps; row of the stacked tables
    // Per-animal terms (fruitfly.py:514-552), combined over the animals of the model: reward terms ADD, termination flags and
    // the three tracking distances take the MAX (any animal off its clip ends the episode).  With one animal (every
    // reference env) this is the reference expression, operation for operation.
    const int NA = m.n_animals;
    float pos_r = 0.f, quat_r = 0.f, joint_r = 0.f, angvel_r = 0.f, bodypos_r = 0.f, endeff_r = 0.f, healthy_r = 0.f;
    float too_far = 0.f, bad_pose = 0.f, bad_quat = 0.f, fall = 0.f, summed = 0.f, quat_d = 0.f, joint_d = 0.f;
    for (int a = 0; a < NA; a++) {
      const Animal an = animal(a);
      const float* qa = qpos() + an.qadr;
      float pd[3] = {0.f, 0.f, 0.f}, qd = 0.f, pr = 0.f, qr = 0.f;
      const int qoff = an.qadr + (m.free_jnt ? 7 : 0);
      if (m.free_jnt) {
        const float* cp = m.clip_position + 3 * (fi * NA + a);
        pd[0] = qa[0] - BT_LDG(cp); pd[1] = qa[1] - BT_LDG(cp + 1); pd[2] = qa[2] - BT_LDG(cp + 2);
        const float sp = (pd[0] + pd[1]) + pd[2];
        pr = m.pos_reward_weight * expf(-400.f * (sp * sp));
        const float* cq = m.clip_quaternion + 4 * (fi * NA + a);
        float qa4[4] = {qa[3], qa[4], qa[5], qa[6]}, b[4] = {BT_LDG(cq), BT_LDG(cq + 1), BT_LDG(cq + 2), BT_LDG(cq + 3)};
        const float na = sqrtf(qa4[0] * qa4[0] + qa4[1] * qa4[1] + qa4[2] * qa4[2] + qa4[3] * qa4[3]);
        const float nb = sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3]);
        float dt = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) dt += (qa4[k] / na) * (b[k] / nb);
        float dd = 2.f * dt * dt - 1.f;
        dd = dd > 1.f ? 1.f : dd;
        const float ang = 0.5f * acosf(dd);
        qd = ang * ang;
        qr = m.quat_reward_weight * expf(-4.0f * qd);
      }
      float js = 0.f;
      for (int k = lane; k < an.nj; k += G) js += qpos()[qoff + k] - BT_LDG(m.clip_joints + (size_t)fi * m.clip_nj + an.jbase + k);
      js = W::allsum(js);
      const float jd = js * js;
      const float* ca = m.clip_angular_velocity + 3 * (fi * NA + a);
      const float* va = qvel() + an.dadr;
      const float av = ((va[3] - BT_LDG(ca)) + (va[4] - BT_LDG(ca + 1))) + (va[5] - BT_LDG(ca + 2));
      float bs = 0.f, es = 0.f;
      for (int it = 3 * BT_LDG(m.bidx_adr + a) + lane, i1 = 3 * BT_LDG(m.bidx_adr + a + 1); it < i1; it += G) {
        const int b = BT_LDG(m.body_idxs + it / 3), k = it % 3;
        bs += xpos()[3 * b + k] - BT_LDG(m.clip_body_positions + ((size_t)fi * m.nbody + b) * 3 + k);
      }
      for (int it = 3 * BT_LDG(m.eidx_adr + a) + lane, i1 = 3 * BT_LDG(m.eidx_adr + a + 1); it < i1; it += G) {
        const int b = BT_LDG(m.endeff_idxs + it / 3), k = it % 3;
        es += xpos()[3 * b + k] - BT_LDG(m.clip_body_positions + ((size_t)fi * m.nbody + b) * 3 + k);
      }
      bs = W::allsum(bs); es = W::allsum(es);
      const float z = xpos()[3 * an.torso + 2];
      float healthy = z < m.healthy_z_min ? 0.f : 1.f;
      healthy = z > m.healthy_z_max ? 0.f : healthy;
      const float w0 = pd[0], w1 = pd[1], w2 = pd[2] * 0.2f;
      const float sm = (w0 * w0 + w1 * w1) + w2 * w2;
      pos_r += pr; quat_r += qr;
      joint_r += m.joint_reward_weight * expf(-0.5f * jd);
      angvel_r += m.angvel_reward_weight * expf(-0.5f * (av * av));
      bodypos_r += m.bodypos_reward_weight * expf(-6.0f * (bs * bs));
      endeff_r += m.endeff_reward_weight * expf(-0.75f * (es * es));
      healthy_r += m.terminate_when_unhealthy ? m.healthy_reward : m.healthy_reward * healthy;
      // max over the animals; a NaN distance propagates (comparisons with NaN are false in fmaxf, so select by hand)
      summed = a == 0 ? sm : (sm > summed || sm != sm ? sm : summed);
      quat_d = a == 0 ? qd : (qd > quat_d || qd != qd ? qd : quat_d);
      joint_d = a == 0 ? jd : (jd > joint_d || jd != jd ? jd : joint_d);
      too_far = fmaxf(too_far, sm > m.too_far_dist ? 1.f : 0.f);
      bad_pose = fmaxf(bad_pose, jd > m.bad_pose_dist ? 1.f : 0.f);
      bad_quat = fmaxf(bad_quat, qd > m.bad_quat_dist ? 1.f : 0.f);
      fall = fmaxf(fall, 1.f - healthy);
    }
    float cs = 0.f;
    for (int u = lane; u < m.nu; u += G) cs += action[u] * action[u];
    cs = W::allsum(cs);
    const float ctrl_cost = m.ctrl_cost_weight * cs;
    float reward = joint_r + pos_r + quat_r + angvel_r + bodypos_r + endeff_r + healthy_r - ctrl_cost;
    float done = m.terminate_when_unhealthy ? fall : 0.f;
    done = fmaxf(fmaxf(done, too_far), fmaxf(bad_pose, bad_quat));
    // NaN guard (fruitfly.py:569-577): any NaN in the pipeline state => done
    int nan = 0;
    for (int i = lane; i < m.nq; i += G) nan |= qpos()[i] != qpos()[i];
    for (int i = lane; i < m.nv; i += G) nan |= (qvel()[i] != qvel()[i]) | (warm()[i] != warm()[i]);
    for (int i = lane; i < m.na; i += G) nan |= act()[i] != act()[i];
    for (int i = lane; i < 3 * m.nbody; i += G) nan |= xpos()[i] != xpos()[i];
    nan |= cs != cs;  // data.ctrl = action is part of the flattened Data too: the sum of squares is NaN iff an action is
    nan = W::any(nan);
    done = fmaxf(done, nan ? 1.f : 0.f);
    r.reward = bt_nan_to_num(reward);
    r.done = done;
    r.metrics[BT_M_POS] = pos_r; r.metrics[BT_M_QUAT] = quat_r; r.metrics[BT_M_JOINT] = joint_r;
    r.metrics[BT_M_ANGVEL] = angvel_r; r.metrics[BT_M_BODYPOS] = bodypos_r; r.metrics[BT_M_ENDEFF] = endeff_r;
    r.metrics[BT_M_QUADCTRL] = -ctrl_cost; r.metrics[BT_M_ALIVE] = healthy_r; r.metrics[BT_M_TOO_FAR] = too_far;
    r.metrics[BT_M_BAD_POSE] = bad_pose; r.metrics[BT_M_BAD_QUAT] = bad_quat; r.metrics[BT_M_FALL] = fall;
    r.summed_pos = summed; r.quat_d = quat_d; r.joint_d = joint_d;
  }
};

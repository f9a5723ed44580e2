"""Packs a compiled model (``mjcf.Model``) + env constants + reference clip into the named flat tables
that the sm_100a kernels read (``csrc/bt_model.h``: the X-macro lists name exactly these arrays).

What the reference does at this point: ``mjx.put_model`` uploads ``mjModel`` and the env keeps the clip
and index lists as jnp arrays (/root/reference/envs/fruitfly.py:405-447).  Here the model is additionally
*re-indexed for warp-per-environment execution*: level schedules for the body tree, tree-sparse mass
matrix layout (MuJoCo's ``dof_Madr`` convention), per-contact ancestor chains instead of a dense ``efc_J``,
and per-dof gather lists so every reduction in the kernels is deterministic (no atomics).

Nothing here runs in the timed path.
"""
from __future__ import annotations

from typing import Dict

import os

import numpy as np

from . import mjcf

# contact functions (csrc/bt_model.h)
FN_PLANE_CAPSULE, FN_PLANE_ELLIPSOID, FN_PLANE_SPHERE, FN_CAPSULE_CAPSULE = 0, 1, 2, 3

CLIP_FIELDS = ("position", "quaternion", "joints", "body_positions", "angular_velocity")


def _i(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32).ravel())


def _f(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).ravel())


def _gather_idx(idx, n):
    """JAX gather semantics of ``x[..., idx]``: negative wraps once, out of range clamps."""
    idx = np.asarray(idx, dtype=np.int64)
    idx = np.where(idx < 0, idx + n, idx)
    return np.clip(idx, 0, n - 1).astype(np.int32)


def pack(m: mjcf.Model, cfg: dict, clip: Dict[str, np.ndarray], lanes: int = 32) -> Dict[str, np.ndarray]:
    """Returns {name: int32/float32 array}; 1-element arrays are the scalars of BtDev."""
    a = m.a
    nq, nv, nu, na, nbody, njnt = m.nq, m.nv, m.nu, m.na, m.nbody, m.njnt
    t: Dict[str, np.ndarray] = {}

    def S(name, v):
        t[name] = np.array([v], dtype=np.int32)

    def SF(name, v):
        t[name] = np.array([v], dtype=np.float32)

    # ------------------------------------------------------------------ body tree
    parent = a["body_parentid"]
    depth = a["body_depth"]
    nlevel = int(depth.max())  # levels 1..max (world = 0 is not scheduled)
    order = sorted(range(1, nbody), key=lambda b: (depth[b], b))
    level_adr = np.zeros(nlevel + 1, dtype=np.int32)
    for b in order:
        level_adr[depth[b]] += 1  # count at index depth (1-based) -> shift below
    counts = [sum(1 for b in order if depth[b] == L) for L in range(1, nlevel + 1)]
    level_adr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    children = [[] for _ in range(nbody)]
    for b in range(1, nbody):
        children[parent[b]].append(b)
    child_adr = np.zeros(nbody + 1, dtype=np.int32)
    child_id = []
    for b in range(nbody):
        child_adr[b] = len(child_id)
        child_id.extend(children[b])
    child_adr[nbody] = len(child_id)
    # body chains (maximal single-child body paths: consecutive ids by DFS numbering) for the pose recursion
    bchain_b0, bchain_len, body_chain = [], [], np.zeros(nbody, dtype=np.int32)
    for b in range(1, nbody):
        pb = parent[b]
        if pb > 0 and len(children[pb]) == 1:
            assert pb == b - 1
            body_chain[b] = body_chain[pb]
            bchain_len[body_chain[b]] += 1
        else:
            body_chain[b] = len(bchain_b0)
            bchain_b0.append(b); bchain_len.append(1)
    nbchain = len(bchain_b0)
    # scheduled by HEIGHT (distance to the deepest leaf chain below), root-most first: a chain only waits for its own
    # ancestors, so short leaf chains hanging off the root run beside the long ones (rodent: 2 + 7 + 30 steps, not 2 + 9 + 30)
    bheight = np.zeros(nbchain, dtype=np.int32)
    for c in reversed(range(nbchain)):   # DFS numbering: children chains have larger ids
        pb = parent[bchain_b0[c]]
        if pb != 0:
            pc = body_chain[pb]
            bheight[pc] = max(bheight[pc], bheight[c] + 1)
    bclevel = bheight.max() - bheight if nbchain else bheight
    nbclev = int(bclevel.max()) + 1
    t["bchain_b0"] = _i(bchain_b0); t["bchain_len"] = _i(bchain_len)
    t["bclev_adr"] = np.concatenate([[0], np.cumsum([int(np.sum(bclevel == L)) for L in range(nbclev)])]).astype(np.int32)
    t["bclev_chain"] = _i(sorted(range(nbchain), key=lambda c: (bclevel[c], c)))
    S("nbchain", nbchain); S("nbclev", nbclev)
    # pointer-jumping ancestor tables of the pose composition: round r composes a body with body_anc[r][b] (its 2^r-th ancestor;
    # 0 = the pose is already a world pose)
    bdepth = np.zeros(nbody, dtype=np.int32)
    for b in range(1, nbody):
        bdepth[b] = bdepth[parent[b]] + 1
    nbanc = max(1, int(np.ceil(np.log2(max(int(bdepth.max()), 1)))))
    anc = np.zeros((nbanc, nbody), dtype=np.int32)
    anc[0] = parent
    for r in range(1, nbanc):
        anc[r] = anc[r - 1][anc[r - 1]]
    assert nbanc == 1 or not anc[nbanc - 1][anc[nbanc - 1]].any()
    S("nbanc", nbanc)
    t["body_anc"] = anc.reshape(-1)
    # per-round work lists of the composition (deepest bodies first, so the lists shrink to fewer 32-lane slots as the
    # rounds go on): a body is composed while its 2^r-th ancestor exists; the round after it reached the world frame it
    # copies its pose to the other ping-pong buffer once (later rounds read it from either).  Item = body | ancestor << 12 |
    # final-after-this-round << 28 | copy-only << 29.
    assert nbody < 4096
    order = sorted(range(1, nbody), key=lambda b: -bdepth[b])
    cmp_adr, cmp_item = [0], []
    for r in range(nbanc):
        for b in order:
            if anc[r][b] > 0:
                fin = 1 if (r == nbanc - 1 or anc[r + 1][b] == 0) else 0
                cmp_item.append(b | (int(anc[r][b]) << 12) | (fin << 28))
        for b in order:   # reached the world frame in the previous round (or, in round 0, directly under the world)
            was_active = r > 0 and anc[r - 1][b] > 0
            if anc[r][b] == 0 and (was_active or r == 0):
                cmp_item.append(b | (1 << 29))
        cmp_adr.append(len(cmp_item))
    t["cmp_adr"] = _i(cmp_adr); t["cmp_item"] = _i(cmp_item)
    # reference point of every kinematic tree: the position of the tree's root body (a free root moves with
    # qpos[0:3]; a static root is a model constant).  MJX uses the subtree COM here; any fixed point gives the
    # same qM / qfrc_bias (DESIGN.md "reference point").
    roots = [b for b in range(1, nbody) if parent[b] == 0]
    root_slot = {b: k for k, b in enumerate(roots)}
    body_ref = np.array([0] + [root_slot[a["body_rootid"][b]] for b in range(1, nbody)], dtype=np.int32)

    S("nq", nq); S("nv", nv); S("nu", nu); S("na", na); S("nbody", nbody); S("njnt", njnt)
    S("nlevel", nlevel); S("nroot", len(roots))
    t["body_parentid"] = _i(parent)
    t["body_jntadr"] = _i(a["body_jntadr"])
    t["body_jntnum"] = _i(a["body_jntnum"])
    t["body_ref"] = body_ref
    # bit 0: body_quat is the identity; bit 1: additionally no joints -> the local frame is a pure translation
    ident = [int(np.abs(a["body_quat"][b] - np.array([1.0, 0, 0, 0])).max() == 0.0) for b in range(nbody)]
    t["body_flags"] = _i([ident[b] | (2 if ident[b] and a["body_jntnum"][b] == 0 else 0) for b in range(nbody)])
    t["body_lastdof"] = _i(a["body_lastdof"])
    t["level_adr"] = level_adr
    t["level_body"] = _i(order)
    t["child_adr"] = child_adr
    t["child_id"] = _i(child_id) if child_id else np.zeros(1, np.int32)
    t["body_pos"] = _f(a["body_pos"]); t["body_quat"] = _f(a["body_quat"])
    t["body_ipos"] = _f(a["body_ipos"]); t["body_iquat"] = _f(a["body_iquat"])
    t["body_mass"] = _f(a["body_mass"]); t["body_inertia"] = _f(a["body_inertia"])
    # fluid (inertia-box) constants: box[3] per body; zero mass -> skipped in the kernel
    box = np.zeros((nbody, 3))
    for b in range(1, nbody):
        mass = a["body_mass"][b]
        if mass > 0:
            I = a["body_inertia"][b]
            bx = np.array([I[1] + I[2] - I[0], I[0] + I[2] - I[1], I[0] + I[1] - I[2]])
            box[b] = np.sqrt(6 * np.maximum(bx, 1e-12) / max(mass, 1e-12))
    t["body_fluidbox"] = _f(box)

    # ------------------------------------------------------------------ joints / dofs
    t["jnt_type"] = _i(a["jnt_type"]); t["jnt_qposadr"] = _i(a["jnt_qposadr"]); t["jnt_dofadr"] = _i(a["jnt_dofadr"])
    t["jnt_pos"] = _f(a["jnt_pos"]); t["jnt_axis"] = _f(a["jnt_axis"])
    # bit 0: the joint sits at the body origin (anchor = body position, no offset rotations)
    t["jnt_flags"] = _i([int(np.abs(a["jnt_pos"][j]).max() == 0.0) for j in range(njnt)])
    t["jnt_bodyid"] = _i(a["jnt_bodyid"])
    for j in range(njnt):
        if a["jnt_type"][j] == mjcf.JNT_FREE:
            bj = a["jnt_bodyid"][j]
            assert parent[bj] == 0 and a["body_jntnum"][bj] == 1, "free joints must be the only joint of a tree root"
    t["qpos0"] = _f(a["qpos0"])
    # packed constant records of body_frame (three 128-bit loads per body / per joint instead of ~10 scalar table reads)
    body_rec = np.zeros((nbody, 12), dtype=np.float32)
    for b in range(nbody):
        body_rec[b, 0:3] = a["body_pos"][b]; body_rec[b, 3:7] = a["body_quat"][b]
        body_rec[b, 7:11] = [a["body_jntadr"][b], a["body_jntnum"][b], parent[b], body_ref[b]]
    t["body_rec"] = body_rec.reshape(-1)
    # body_local's constants: ipos iquat inertia mass | fluidbox lastdof ref (16 floats)
    bl_rec = np.zeros((nbody, 16), dtype=np.float32)
    for b in range(nbody):
        bl_rec[b, 0:3] = a["body_ipos"][b]; bl_rec[b, 3:7] = a["body_iquat"][b]; bl_rec[b, 7:10] = a["body_inertia"][b]
        bl_rec[b, 10] = a["body_mass"][b]
        bl_rec[b, 11:14] = box[b]; bl_rec[b, 14] = a["body_lastdof"][b]; bl_rec[b, 15] = body_ref[b]
    t["bl_rec"] = bl_rec.reshape(-1)
    jnt_rec = np.zeros((max(njnt, 1), 12), dtype=np.float32)
    for j in range(njnt):
        qa = int(a["jnt_qposadr"][j])
        jnt_rec[j, 0:4] = [a["jnt_type"][j], qa, a["jnt_dofadr"][j], float(np.abs(a["jnt_pos"][j]).max() == 0.0)]
        jnt_rec[j, 4:7] = a["jnt_pos"][j]; jnt_rec[j, 7:10] = a["jnt_axis"][j]
        jnt_rec[j, 10] = a["qpos0"][qa]
    t["jnt_rec"] = jnt_rec.reshape(-1)
    dof_jnt = a["dof_jntid"]
    dof_qadr = np.full(nv, -1, dtype=np.int32)
    dof_stiff = np.zeros(nv); dof_spring = np.zeros(nv)
    dof_lim = np.zeros(nv, dtype=np.int32)
    dof_range = np.zeros((nv, 2)); dof_solref = np.zeros((nv, 2)); dof_solimp = np.zeros((nv, 5))
    dof_margin = np.zeros(nv)
    for i in range(nv):
        j = dof_jnt[i]
        if a["jnt_type"][j] == mjcf.JNT_HINGE:
            qa = a["jnt_qposadr"][j]
            dof_qadr[i] = qa
            dof_stiff[i] = a["jnt_stiffness"][j]
            dof_spring[i] = a["qpos_spring"][qa]
            dof_lim[i] = int(a["jnt_limited"][j])
            dof_range[i] = a["jnt_range"][j]
            dof_solref[i] = a["jnt_solref"][j]
            dof_solimp[i] = a["jnt_solimp"][j]
            dof_margin[i] = a["jnt_margin"][j]
    t["dof_bodyid"] = _i(a["dof_bodyid"]); t["dof_parentid"] = _i(a["dof_parentid"])
    # velocity-sweep flag: 0 hinge, 1 free translation, 2 first / 3 later rotational dof of a free joint
    vflag = np.zeros(nv, dtype=np.int32)
    for j in range(njnt):
        if a["jnt_type"][j] == mjcf.JNT_FREE:
            d0 = a["jnt_dofadr"][j]
            vflag[d0:d0 + 3] = 1; vflag[d0 + 3] = 2; vflag[d0 + 4:d0 + 6] = 3
    t["dof_vflag"] = vflag
    t["dof_qposadr"] = dof_qadr; t["dof_limited"] = dof_lim
    t["dof_stiffness"] = _f(dof_stiff); t["dof_springref"] = _f(dof_spring)
    t["dof_armature"] = _f(a["dof_armature"]); t["dof_damping"] = _f(a["dof_damping"])
    t["dof_range"] = _f(dof_range); t["dof_solref"] = _f(dof_solref); t["dof_solimp"] = _f(dof_solimp)
    t["dof_margin"] = _f(dof_margin); t["dof_invweight0"] = _f(a["dof_invweight0"])
    # dof tree cut into chains (maximal single-child paths: consecutive dof ids by DFS numbering), scheduled by the depth
    # of the chain in the chain tree (articulated-body sweeps, csrc/bt_impl.h::aba_factor / solve / mul_M)
    dchild = [[] for _ in range(nv)]
    for i in range(nv):
        if a["dof_parentid"][i] >= 0:
            dchild[a["dof_parentid"][i]].append(i)
    dchild_adr, dchild_id = [0], []
    for i in range(nv):
        dchild_id.extend(dchild[i])
        dchild_adr.append(len(dchild_id))
    # bodies carried by a dof: every body whose last dof on the chain root -> body is that dof (jointless bodies ride
    # on their ancestor's last dof; bodies of the static world group carry no dof and drop out of the dynamics)
    dofbody = [[] for _ in range(nv)]
    for b in range(1, nbody):
        if a["body_lastdof"][b] >= 0:
            dofbody[a["body_lastdof"][b]].append(b)
    dofbody_adr, dofbody_id = [0], []
    for i in range(nv):
        dofbody_id.extend(dofbody[i])
        dofbody_adr.append(len(dofbody_id))
    chain_k0, chain_len, dof_chain = [], [], np.zeros(nv, dtype=np.int32)
    for i in range(nv):
        par = a["dof_parentid"][i]
        if par >= 0 and len(dchild[par]) == 1:
            assert par == i - 1, "single-child dofs must be consecutive (DFS numbering)"
            dof_chain[i] = dof_chain[par]
            chain_len[dof_chain[i]] += 1
        else:
            dof_chain[i] = len(chain_k0)
            chain_k0.append(i); chain_len.append(1)
    nchain = len(chain_k0)
    clevel = np.zeros(nchain, dtype=np.int32)
    for c in range(nchain):
        par = a["dof_parentid"][chain_k0[c]]
        clevel[c] = 0 if par < 0 else clevel[dof_chain[par]] + 1
    nclev = int(clevel.max()) + 1 if nchain else 1
    corder = sorted(range(nchain), key=lambda c: (clevel[c], c))
    clev_adr = np.concatenate([[0], np.cumsum([int(np.sum(clevel == L)) for L in range(nclev)])]).astype(np.int32)
    S("nchain", nchain); S("nclev", nclev)
    t["chain_k0"] = _i(chain_k0); t["chain_len"] = _i(chain_len); t["clev_adr"] = clev_adr; t["clev_chain"] = _i(corder)
    t["dof_chain"] = dof_chain
    # the one-lane-per-chain sweeps are scheduled by chain HEIGHT instead (see the body chains above); aba_factor keeps the
    # depth levels, which pack its four 8-lane groups better
    cheight = np.zeros(nchain, dtype=np.int32)
    for c in reversed(range(nchain)):
        par = a["dof_parentid"][chain_k0[c]]
        if par >= 0:
            pc = dof_chain[par]
            cheight[pc] = max(cheight[pc], cheight[c] + 1)
    hlevel = cheight.max() - cheight if nchain else cheight
    nhlev = int(hlevel.max()) + 1 if nchain else 1
    S("nhlev", nhlev)
    t["hlev_adr"] = np.concatenate([[0], np.cumsum([int(np.sum(hlevel == L)) for L in range(nhlev)])]).astype(np.int32)
    t["hlev_chain"] = _i(sorted(range(nchain), key=lambda c: (hlevel[c], c)))
    t["dchild_adr"] = _i(dchild_adr); t["dchild_id"] = _i(dchild_id) if dchild_id else np.zeros(1, np.int32)
    t["dofbody_adr"] = _i(dofbody_adr); t["dofbody_id"] = _i(dofbody_id) if dofbody_id else np.zeros(1, np.int32)
    # link records: the inertia / RNE force of all bodies carried by a dof are summed into the slot of its FIRST body
    # (tree_forward's merge pass), so the serial sweeps read one record per dof: dof_irec = that body, or -1
    dof_irec = np.full(nv, -1, dtype=np.int32)
    merge_adr, merge_dst, merge_src = [0], [], []
    for i in range(nv):
        if dofbody[i]:
            dof_irec[i] = dofbody[i][0]
            if len(dofbody[i]) > 1:
                merge_dst.append(dofbody[i][0]); merge_src.extend(dofbody[i][1:]); merge_adr.append(len(merge_src))
    S("nmerge", len(merge_dst))
    t["dof_irec"] = dof_irec; t["merge_adr"] = _i(merge_adr)
    t["merge_dst"] = _i(merge_dst) if merge_dst else np.zeros(1, np.int32)
    t["merge_src"] = _i(merge_src) if merge_src else np.zeros(1, np.int32)

    def chain(body):
        out, d = [], a["body_lastdof"][body]
        while d >= 0:
            out.append(int(d))
            d = a["dof_parentid"][d]
        return out

    # ------------------------------------------------------------------ contacts (static list)
    con = []  # one record per potential contact
    cgeoms, cgeom_slot = [], {}

    def cg(g):
        if g not in cgeom_slot:
            cgeom_slot[g] = len(cgeoms)
            cgeoms.append(g)
        return cgeom_slot[g]

    for p in range(len(a["pair_ncon"])):
        g1, g2 = int(a["pair_geom"][p, 0]), int(a["pair_geom"][p, 1])
        t1, t2 = a["geom_type"][g1], a["geom_type"][g2]
        if (t1, t2) == (mjcf.GEOM_PLANE, mjcf.GEOM_CAPSULE):
            fn = FN_PLANE_CAPSULE
        elif (t1, t2) == (mjcf.GEOM_PLANE, mjcf.GEOM_ELLIPSOID):
            fn = FN_PLANE_ELLIPSOID
        elif (t1, t2) == (mjcf.GEOM_PLANE, mjcf.GEOM_SPHERE):
            fn = FN_PLANE_SPHERE
        elif (t1, t2) == (mjcf.GEOM_CAPSULE, mjcf.GEOM_CAPSULE):
            fn = FN_CAPSULE_CAPSULE
        else:
            raise NotImplementedError((t1, t2))
        b1, b2 = int(a["geom_bodyid"][g1]), int(a["geom_bodyid"][g2])
        dim = int(a["pair_condim"][p])
        if dim not in (1, 3):
            raise NotImplementedError("condim 4/6 is unused by the selected assets (MJX 3.2 does not support it)")
        for sub in range(int(a["pair_ncon"][p])):
            con.append(dict(g1=cg(g1), g2=cg(g2), b1=b1, b2=b2, fn=fn, sub=sub, dim=dim,
                            mu=(a["pair_friction"][p, 0], a["pair_friction"][p, 1]),
                            solref=a["pair_solref"][p], solimp=a["pair_solimp"][p],
                            includemargin=a["pair_margin"][p] - a["pair_gap"][p],
                            invw=a["body_invweight0"][b1, 0] + a["body_invweight0"][b2, 0]))
    ncon = len(con)
    S("ncon", ncon)
    S("ncgeom", len(cgeoms))
    t["cgeom_bodyid"] = _i([a["geom_bodyid"][g] for g in cgeoms]) if cgeoms else np.zeros(1, np.int32)
    t["cgeom_pos"] = _f([a["geom_pos"][g] for g in cgeoms]) if cgeoms else np.zeros(3, np.float32)
    t["cgeom_quat"] = _f([a["geom_quat"][g] for g in cgeoms]) if cgeoms else np.zeros(4, np.float32)
    t["cgeom_size"] = _f([a["geom_size"][g] for g in cgeoms]) if cgeoms else np.zeros(3, np.float32)
    # contact bodies (bodies that carry dofs and appear in a contact): chain lists for the matrix-free J
    # keyed by the body's last dof: bodies riding on the same dof share one slot (same chain, same sums)
    cbs, cb_slot, dof_slot = [], {}, {}
    for c in con:
        for b in (c["b1"], c["b2"]):
            d = int(a["body_lastdof"][b])
            if d >= 0 and b not in cb_slot:
                if d not in dof_slot:
                    dof_slot[d] = len(cbs)
                    cbs.append(b)
                cb_slot[b] = dof_slot[d]
    ncb = len(cbs)
    S("ncb", ncb)
    t["cb_lastdof"] = _i([a["body_lastdof"][b] for b in cbs]) if cbs else np.zeros(1, np.int32)
    # the root->leaves sweeps walk a dof chain in segments that end at the last dof of a contact body (where the running
    # spatial acceleration is the chain sum the constraint Jacobian needs) or at the chain end (seg_cb = -1)
    seg_adr, seg_end, seg_cb = [0], [], []
    for c in range(nchain):
        kb = chain_k0[c] + chain_len[c] - 1
        for d in range(chain_k0[c], kb + 1):
            if d in dof_slot:
                seg_end.append(d); seg_cb.append(dof_slot[d])
        if kb not in dof_slot:
            seg_end.append(kb); seg_cb.append(-1)
        seg_adr.append(len(seg_end))
    t["seg_adr"] = _i(seg_adr); t["seg_end"] = _i(seg_end); t["seg_cb"] = _i(seg_cb)
    # one 8-int descriptor per chain (two 128-bit loads in the sweep prologues): k0, kb, parent chain, first child slot,
    # number of child chains, first segment, number of segments, parent dof
    cchild = [[] for _ in range(nchain)]
    cparent = np.full(nchain, -1, dtype=np.int32)
    for c in range(nchain):
        par = a["dof_parentid"][chain_k0[c]]
        if par >= 0:
            cparent[c] = dof_chain[par]
            assert par == chain_k0[cparent[c]] + chain_len[cparent[c]] - 1, "child chains attach at the end of the parent chain"
            cchild[cparent[c]].append(c)
    cchild_adr, cchild_id = [0], []
    for c in range(nchain):
        cchild_id.extend(cchild[c]); cchild_adr.append(len(cchild_id))
    # k0, kb, parent chain, chain id | first child slot, #children + (#segments << 16), first segment, parent dof
    desc = np.zeros((max(nchain, 1), 8), dtype=np.int32)
    for c in range(nchain):
        desc[c] = [chain_k0[c], chain_k0[c] + chain_len[c] - 1, cparent[c], c, cchild_adr[c],
                   len(cchild[c]) | ((seg_adr[c + 1] - seg_adr[c]) << 16), seg_adr[c], a["dof_parentid"][chain_k0[c]]]
    t["chain_desc"] = desc.reshape(-1)
    # the one-lane-per-chain sweeps read their chain straight from a per-(pass, lane) copy of the descriptor: a pass is up to
    # 32 chains of one height level (root-most level first); empty lanes carry kb < k0
    passes = []
    for L in range(nhlev):
        cs = [c for c in sorted(range(nchain), key=lambda c: (hlevel[c], c)) if hlevel[c] == L]
        for q in range(0, len(cs), 32):
            row = np.zeros((32, 8), dtype=np.int32)
            row[:, 1] = -1
            for ln, c in enumerate(cs[q:q + 32]):
                row[ln] = desc[c]
            passes.append(row)
    # the factor sweep walks four chains at a time (8-lane groups), leaves first; a pass lasts as long as its longest chain.
    # List scheduling by critical path: among the chains whose children are all done, a pass takes the four with the longest
    # remaining way to the root (own dofs + ancestors'), so the chains nobody waits for fill the groups that would idle
    # (rodent: 24 + 8 + 6 = 38 serial dof steps instead of 24 + 9 + 6 by depth level; two rodents 53 instead of 72)
    to_root = np.zeros(nchain, dtype=np.int64)
    for c in sorted(range(nchain), key=lambda c: clevel[c]):
        to_root[c] = chain_len[c] + (to_root[cparent[c]] if cparent[c] >= 0 else 0)
    apasses, done_c = [], set()
    while len(done_c) < nchain:
        ready = [c for c in range(nchain) if c not in done_c and all(ch in done_c for ch in cchild[c])]
        take = sorted(ready, key=lambda c: (-to_root[c], c))[:4]
        row = np.zeros((4, 8), dtype=np.int32)
        row[:, 1] = -1
        for g, c in enumerate(take):
            row[g] = desc[c]
        apasses.append(row)
        done_c.update(take)
    S("napass", len(apasses))
    t["apass_desc"] = np.concatenate(apasses).reshape(-1) if apasses else np.zeros(32, np.int32)
    S("nhpass", len(passes))
    t["hpass_desc"] = np.concatenate(passes).reshape(-1) if passes else np.zeros(256, np.int32)
    t["cchild_id"] = _i(cchild_id) if cchild_id else np.zeros(1, np.int32)
    cb_adr, cb_dof = [0], []
    for b in cbs:
        cb_dof.extend(chain(b))
        cb_adr.append(len(cb_dof))
    t["cb_adr"] = _i(cb_adr)
    t["cb_dof"] = _i(cb_dof) if cb_dof else np.zeros(1, np.int32)
    t["cb_ref"] = _i([body_ref[b] for b in cbs]) if cbs else np.zeros(1, np.int32)

    def Z(n, dt):
        return np.zeros(max(n, 1), dtype=dt)

    t["con_g1"] = _i([c["g1"] for c in con]) if con else Z(1, np.int32)
    t["con_g2"] = _i([c["g2"] for c in con]) if con else Z(1, np.int32)
    t["con_cb1"] = _i([cb_slot.get(c["b1"], -1) for c in con]) if con else Z(1, np.int32)
    t["con_cb2"] = _i([cb_slot.get(c["b2"], -1) for c in con]) if con else Z(1, np.int32)
    t["con_ref"] = _i([body_ref[c["b2"]] if a["body_lastdof"][c["b2"]] >= 0 else body_ref[c["b1"]] for c in con]) if con else Z(1, np.int32)
    t["con_fn"] = _i([c["fn"] for c in con]) if con else Z(1, np.int32)
    t["con_sub"] = _i([c["sub"] for c in con]) if con else Z(1, np.int32)
    t["con_dim"] = _i([c["dim"] for c in con]) if con else Z(1, np.int32)
    t["con_mu"] = _f([c["mu"] for c in con]) if con else Z(2, np.float32)
    t["con_solref"] = _f([c["solref"] for c in con]) if con else Z(2, np.float32)
    t["con_solimp"] = _f([c["solimp"] for c in con]) if con else Z(5, np.float32)
    t["con_includemargin"] = _f([c["includemargin"] for c in con]) if con else Z(1, np.float32)
    t["con_invweight"] = _f([c["invw"] for c in con]) if con else Z(1, np.float32)
    # A contact between two kinematic trees (BASELINE.json configs[3]: the opened inter-animal pairs): each tree's spatial
    # quantities are about its own reference point, so body 1's side of such a contact uses the offset from ITS reference
    # (con_xref = tree of body 1, -1 for ordinary contacts) and, in J' f, a second wrench slot behind the ncon regular ones
    # (con_xslot) holding the same force as a wrench about that point.  con_ref stays the tree of body 2.
    con_xref, con_xslot, ncross = [], [], 0
    for c in con:
        m1 = a["body_lastdof"][c["b1"]] >= 0
        m2 = a["body_lastdof"][c["b2"]] >= 0
        if m1 and m2 and body_ref[c["b1"]] != body_ref[c["b2"]]:
            con_xref.append(int(body_ref[c["b1"]])); con_xslot.append(ncon + ncross); ncross += 1
        else:
            con_xref.append(-1); con_xslot.append(-1)
    S("ncross", ncross)
    t["con_xref"] = _i(con_xref) if con else Z(1, np.int32)
    t["con_xslot"] = _i(con_xslot) if con else Z(1, np.int32)
    # J' f in two gathers: contact -> contact body (sign +1 when the body is geom2's, -1 when it is geom1's), then
    # contact body -> every dof on its ancestor chain (common ancestors of a two-body contact cancel in the sum)
    cbcon = [[] for _ in range(max(ncb, 1))]
    for ci, c in enumerate(con):
        if c["b2"] in cb_slot:
            cbcon[cb_slot[c["b2"]]].append((ci, 1.0))
        if c["b1"] in cb_slot:
            cbcon[cb_slot[c["b1"]]].append((con_xslot[ci] if con_xslot[ci] >= 0 else ci, -1.0))
    cbcon_adr, cbcon_c, cbcon_s = [0], [], []
    for k in range(ncb):
        for ci, sg in cbcon[k]:
            cbcon_c.append(ci); cbcon_s.append(sg)
        cbcon_adr.append(len(cbcon_c))
    # contact -> contact-body sums by a segmented warp-shuffle reduction (csrc/bt_impl.h::jt_force) when every contact has exactly
    # one moving body (geom2's), the contacts of a contact body are consecutive, and all contacts fit one lane each
    con_cb = [cb_slot.get(c["b2"], -1) for c in con]
    jt_seg = (0 < ncon <= lanes and ncross == 0 and all(c["b1"] not in cb_slot for c in con) and all(x >= 0 for x in con_cb)
              and all(con_cb[i] <= con_cb[i + 1] for i in range(ncon - 1)))
    seg = np.zeros(max(ncon, 1), dtype=np.int32)
    steps = 0
    if jt_seg:
        maxlen = max(con_cb.count(k) for k in set(con_cb))
        steps = int(np.ceil(np.log2(maxlen))) if maxlen > 1 else 0
        for i in range(ncon):
            bits = 0
            for st in range(steps):
                j = i + (1 << st)
                if j < ncon and con_cb[j] == con_cb[i]:
                    bits |= 1 << st
            head = con_cb[i] + 1 if (i == 0 or con_cb[i - 1] != con_cb[i]) else 0
            seg[i] = bits | (head << 8)
        steps = max(steps, 1)      # 0 would disable the path: a model whose segments all have one contact still takes it
    S("jt_seg_steps", steps if jt_seg else 0)
    t["con_seg"] = seg
    dofcb = [[] for _ in range(nv)]
    for k, b in enumerate(cbs):
        for d in chain(b):
            dofcb[d].append(k)
    dofcb_adr, dofcb_id = [0], []
    for d in range(nv):
        dofcb_id.extend(dofcb[d])
        dofcb_adr.append(len(dofcb_id))
    # dofs with the same set of contact bodies below them share one summed wrench: J' f per dof is then ONE dot product
    # (the root dofs of the rodent see all 8 contact bodies)
    grp_of, grp_list = {}, []
    dof_wgrp = np.full(nv, -1, dtype=np.int32)
    for d in range(nv):
        if dofcb[d]:
            key = tuple(sorted(dofcb[d]))
            if key not in grp_of:
                grp_of[key] = len(grp_list); grp_list.append(key)
            dof_wgrp[d] = grp_of[key]
    wgrp_adr, wgrp_cb = [0], []
    for key in grp_list:
        wgrp_cb.extend(key); wgrp_adr.append(len(wgrp_cb))
    S("nwgrp", len(grp_list))
    # contact bodies are numbered in DFS order, so the ones below a dof are a contiguous range: first | count << 16 (one load, no
    # dependent index loads in the summation loop); 0 in wgrp_contig keeps the index list
    contig = all(list(key) == list(range(key[0], key[0] + len(key))) for key in grp_list) and os.environ.get("BT_WGRP_CONTIG", "1") != "0"
    S("wgrp_contig", int(contig))
    t["wgrp_rng"] = _i([key[0] | (len(key) << 16) for key in grp_list]) if grp_list and contig else Z(max(len(grp_list), 1), np.int32)
    assert 6 * len(grp_list) <= 6 * max(ncon, 1), "the group wrench sums reuse the per-contact wrench slots"
    t["dof_wgrp"] = dof_wgrp; t["wgrp_adr"] = _i(wgrp_adr); t["wgrp_cb"] = _i(wgrp_cb) if wgrp_cb else Z(1, np.int32)
    t["cbcon_adr"] = _i(cbcon_adr); t["cbcon_c"] = _i(cbcon_c) if cbcon_c else Z(1, np.int32)
    # contact index with the sign in the top bit (one load per term of the contact-body wrench sums)
    t["cbcon_cs"] = (np.array([c | (0x80000000 if sg < 0 else 0) for c, sg in zip(cbcon_c, cbcon_s)], dtype=np.uint32).view(np.int32)
                     if cbcon_c else Z(1, np.int32))
    t["cbcon_sign"] = _f(cbcon_s) if cbcon_s else Z(1, np.float32)
    t["dofcb_adr"] = _i(dofcb_adr); t["dofcb_id"] = _i(dofcb_id) if dofcb_id else Z(1, np.int32)

    # ------------------------------------------------------------------ actuators
    wrap_adr, wrap_q, wrap_coef = [0], [], []
    dofact = [[] for _ in range(nv)]
    for u in range(nu):
        gear = a["actuator_gear"][u]
        if a["actuator_trntype"][u] == mjcf.TRN_JOINT:
            j = a["actuator_trnid"][u]
            if a["jnt_type"][j] != mjcf.JNT_HINGE:
                raise NotImplementedError("actuators on free joints are unused")
            wrap_q.append(a["jnt_qposadr"][j]); wrap_coef.append(1.0)
            dofact[a["jnt_dofadr"][j]].append((u, gear))
        else:
            tt = a["actuator_trnid"][u]
            for w in range(a["tendon_adr"][tt], a["tendon_adr"][tt] + a["tendon_num"][tt]):
                j = a["wrap_jntid"][w]
                wrap_q.append(a["jnt_qposadr"][j]); wrap_coef.append(a["wrap_coef"][w])
                dofact[a["jnt_dofadr"][j]].append((u, a["wrap_coef"][w] * gear))
        wrap_adr.append(len(wrap_q))
    t["act_wrap_adr"] = _i(wrap_adr)
    t["act_wrap_qadr"] = _i(wrap_q) if wrap_q else Z(1, np.int32)
    t["act_wrap_coef"] = _f(wrap_coef) if wrap_coef else Z(1, np.float32)
    # dof index of every wrap entry (hinge: qposadr -> dofadr)
    q2d = {int(a["jnt_qposadr"][j]): int(a["jnt_dofadr"][j]) for j in range(njnt) if a["jnt_type"][j] == mjcf.JNT_HINGE}
    t["act_wrap_dadr"] = _i([q2d[int(q)] for q in wrap_q]) if wrap_q else Z(1, np.int32)
    dofact_adr, dofact_u, dofact_coef = [0], [], []
    for d in range(nv):
        for u, cf in dofact[d]:
            dofact_u.append(u); dofact_coef.append(cf)
        dofact_adr.append(len(dofact_u))
    # packed per-actuator / per-dof constant records of smooth_forces (128-bit loads instead of ~20 scalar table reads)
    BIG = np.float32(3.0e38)
    act_rec = np.zeros((max(nu, 1), 16), dtype=np.float32)
    for u in range(nu):
        lim = bool(a["actuator_ctrllimited"][u]); flim = bool(a["actuator_forcelimited"][u])
        aff_g = int(a["actuator_gaintype"][u]) == 1; aff_b = int(a["actuator_biastype"][u]) == 1
        gp, bp = a["actuator_gainprm"][u], a["actuator_biasprm"][u]
        act_rec[u] = [a["actuator_gear"][u],
                      a["actuator_ctrlrange"][u][0] if lim else -BIG, a["actuator_ctrlrange"][u][1] if lim else BIG,
                      max(float(a["actuator_dynprm"][u][0]), 1e-15),
                      gp[0], gp[1] if aff_g else 0.0, gp[2] if aff_g else 0.0, 0.0,
                      bp[0] if aff_b else 0.0, bp[1] if aff_b else 0.0, bp[2] if aff_b else 0.0, float(a["actuator_actadr"][u]),
                      a["actuator_forcerange"][u][0] if flim else -BIG, a["actuator_forcerange"][u][1] if flim else BIG,
                      float(wrap_adr[u]), float(wrap_adr[u + 1] - wrap_adr[u])]
    t["act_rec"] = act_rec.reshape(-1)
    wrap_d = [q2d[int(q)] for q in wrap_q]
    t["wrap_rec"] = (np.array([[cf, q, d, 0.0] for cf, q, d in zip(wrap_coef, wrap_q, wrap_d)], dtype=np.float32).reshape(-1)
                     if wrap_q else Z(4, np.float32))
    dof_rec = np.zeros((nv, 8), dtype=np.float32)
    for d in range(nv):
        dof_rec[d] = [a["dof_damping"][d], dof_stiff[d], dof_spring[d], float(dof_qadr[d]), float(dofact_adr[d]),
                      float(dofact_adr[d + 1] - dofact_adr[d]), 0.0, 0.0]
    t["dof_rec"] = dof_rec.reshape(-1)
    t["dofact_rec"] = (np.array([[cf, u] for cf, u in zip(dofact_coef, dofact_u)], dtype=np.float32).reshape(-1)
                       if dofact_u else Z(2, np.float32))
    t["dofact_adr"] = _i(dofact_adr)
    t["dofact_u"] = _i(dofact_u) if dofact_u else Z(1, np.int32)
    t["dofact_coef"] = _f(dofact_coef) if dofact_coef else Z(1, np.float32)
    for k in ("dyntype", "gaintype", "biastype", "ctrllimited", "forcelimited", "actadr"):
        t["actuator_" + k] = _i(a["actuator_" + k]) if nu else Z(1, np.int32)
    for k in ("gear", "gainprm", "biasprm", "dynprm", "ctrlrange", "forcerange"):
        t["actuator_" + k] = _f(a["actuator_" + k]) if nu else Z(3, np.float32)

    # ------------------------------------------------------------------ options
    S("cone", m.cone); S("iterations", m.iterations); S("ls_iterations", m.ls_iterations)
    S("n_frames", cfg["n_frames"])
    # phase alignment of the kernels' warps (bits, csrc/bt_impl.h::substep); the default is chosen at the end of pack(), once the
    # environments per SM are known
    # initcheck substitute: fill the scratch slice with NaN before every program (csrc/bt_impl.h::poison_scratch)
    S("poison", int(os.environ.get("BT_POISON", "0")))
    SF("timestep", m.timestep)
    SF("grav_x", m.gravity[0]); SF("grav_y", m.gravity[1]); SF("grav_z", m.gravity[2])
    SF("density", m.density); SF("viscosity", m.viscosity); SF("impratio", m.impratio)
    SF("tolerance", m.tolerance); SF("ls_tolerance", m.ls_tolerance); SF("meaninertia", m.meaninertia)

    # ------------------------------------------------------------------ env layer (fruitfly.py:405-447)
    # one clip `[T, ...]`, or several stacked on a leading clip axis `[C, T, ...]` (preprocess.py:254-258: RodentMultiClip)
    multi = np.asarray(clip["joints"]).ndim == 3
    NC = int(np.asarray(clip["joints"]).shape[0]) if multi else 1
    clip = {k: (np.asarray(v) if multi else np.asarray(v)[None]) for k, v in clip.items() if k in CLIP_FIELDS}
    T = int(clip["joints"].shape[1])
    nj = int(clip["joints"].shape[2])
    S("n_clips", NC)
    S("free_jnt", int(cfg["free_jnt"])); S("seed_root_from_clip", int(cfg["seed_root_from_clip"]))
    S("ref_len", cfg["ref_len"]); S("clip_len", T); S("clip_nj", nj)
    # per-animal index lists (configs.resolve: cfg["animals"]); joint ids index the ANIMAL's own joint columns
    animals = cfg.get("animals") or [dict(qadr=0, dadr=0, nj=nj, jbase=0, torso_idx=cfg["torso_idx"], joint_idxs=cfg["joint_idxs"],
                                          body_idxs=cfg["body_idxs"], endeff_idxs=cfg["endeff_idxs"])]
    NA = len(animals)
    S("n_animals", NA)
    jidx, bidx, eidx, jadr, badr, eadr = [], [], [], [0], [0], [0]
    arec = np.zeros((NA, 8), dtype=np.int32)
    for k, an in enumerate(animals):
        jidx.extend(_gather_idx(an["joint_idxs"], an["nj"])); jadr.append(len(jidx))
        bidx.extend(_gather_idx(an["body_idxs"], nbody)); badr.append(len(bidx))
        eidx.extend(_gather_idx(an["endeff_idxs"], nbody)); eadr.append(len(eidx))
        arec[k, :5] = [an["qadr"], an["dadr"], an["nj"], an["jbase"], int(an["torso_idx"]) % nbody]
    jidx, bidx, eidx = _i(jidx), _i(bidx), _i(eidx)
    t["animal_rec"] = arec.reshape(-1); t["jidx_adr"] = _i(jadr); t["bidx_adr"] = _i(badr); t["eidx_adr"] = _i(eadr)
    assert sum(an["nj"] for an in animals) == nj, "the clip's joint columns are the animals' joints side by side"
    S("n_joint_idxs", len(jidx)); S("n_body_idxs", len(bidx)); S("n_endeff_idxs", len(eidx))
    t["joint_idxs"] = jidx; t["body_idxs"] = bidx; t["endeff_idxs"] = eidx if len(eidx) else Z(1, np.int32)
    S("torso_idx", int(cfg["torso_idx"]) % nbody)  # negative ids index from the end, as in JAX
    S("terminate_when_unhealthy", int(cfg["terminate_when_unhealthy"]))
    sfc = float(cfg["steps_for_cur_frame"])
    # fruitfly.py:504-509 compares an int32 counter with this float; a non-integral value never matches
    S("steps_for_cur_frame", int(sfc) if sfc == int(sfc) else -1)
    S("episode_length", cfg["episode_length"]); S("start_frame_range", cfg.get("start_frame_range", 44))
    for k in ("too_far_dist", "bad_pose_dist", "bad_quat_dist", "ctrl_cost_weight", "pos_reward_weight",
              "quat_reward_weight", "joint_reward_weight", "angvel_reward_weight", "bodypos_reward_weight",
              "endeff_reward_weight", "healthy_reward", "reset_noise_scale"):
        SF(k, cfg[k])
    SF("healthy_z_min", cfg["healthy_z_range"][0]); SF("healthy_z_max", cfg["healthy_z_range"][1])
    L = int(cfg["ref_len"])
    if cfg["free_jnt"]:
        obs_size = nq + nv + NA * (3 * L + 4 * L) + len(jidx) * L + 3 * len(bidx) * L
    else:
        obs_size = nq + nv + len(jidx) * L + 3 * len(bidx) * L
    S("obs_size", obs_size)
    for k in CLIP_FIELDS:
        t["clip_" + k] = _f(clip[k])
    assert clip["body_positions"].shape == (NC, T, nbody, 3), clip["body_positions"].shape
    assert nj == (nq - 7 * NA if cfg["free_jnt"] else nq), (nj, nq)
    for k, w in (("position", 3), ("quaternion", 4), ("angular_velocity", 3)):
        assert clip[k].reshape(NC, T, -1).shape[2] == w * NA, (k, clip[k].shape, NA)

    # ------------------------------------------------------------------ per-environment scratch layout (floats)
    lay, off = {}, 0

    def R(name, n):
        nonlocal off
        lay[name] = off
        off += int(n)

    def ALIGN4():
        nonlocal off
        off += (-off) % 4

    R("qpos", nq); R("qvel", nv); R("act", max(na, 1)); R("ctrl", max(nu, 1)); R("warm", nv)
    R("xpos", 3 * nbody)
    ALIGN4()   # quaternions move as one 128-bit access
    R("xquat", 4 * nbody)
    ALIGN4()
    R("cdof", 12 * nv)   # per dof: S_k = cdof_k (6) | G_k = U_k / D_k (6); 16-byte aligned records
    R("crb", 10 * nbody); R("Dinv", nv); R("Dd", nv)
    R("cbJ", 18 * max(ncb, 1))   # three sets of per-contact-body chain sums (qvel, qacc_warmstart, qacc_smooth)
    # T region: cfrc (tree passes) -> the 6x6 reduced articulated inertia of every chain top (aba_factor)
    # -> contact geometry + wrenches + chain sums (solver)
    ALIGN4()   # the second pose buffer's quaternions (T + round4(3 * nbody)) are 128-bit accesses too
    # The per-contact wrenches of J' f (6 per contact, + 6 per inter-tree contact) live only inside the CG loop, where the pivots
    # Dd (used by M v before the solve) and the three cbJ chain-sum sets (consumed by the warm-start selection) are dead: when
    # they fit, the wrenches take that space instead of lengthening T (two-rodent model: 6 instead of 5 environments per SM)
    wrench_floats = 6 * (ncon + ncross)
    wrench_alias = wrench_floats <= nv + 18 * max(ncb, 1)
    R("T", max(7 * nbody + 3, 6 * nv, 12 * ncon + (0 if wrench_alias else wrench_floats) + 6 * max(ncb, 1)))   # 7 * nbody: second pose buffer
    R("ref", 3 * max(len(roots), 1))
    R("actdot", max(na, 1))
    # pvec (sweep state, 6/dof) and the solver vectors behind it are contiguous: together they hold cvel/cacc (12/dof)
    # during the forward tree pass, when none of them is live
    ALIGN4()   # 16-byte aligned: the 12-float (cvel | cacc) records of the velocity sweep move with 128-bit accesses
    w12 = off
    R("pvec", max(6 * nv, 48 * nchain))  # also the 8 x 6 chain-top rows of aba_factor
    for v in ("qfrc_smooth", "qacc_smooth", "qacc", "search", "qfrc_c"):
        R(v, nv)
    R("x", max(nv, nu))
    if off - w12 < 12 * nv:    # inclusive cvel / cacc per dof during the velocity sweep
        off = w12 + 12 * nv
    assert lay["cbJ"] == lay["Dd"] + nv
    lay["wrench"] = lay["Dd"] if wrench_alias else lay["T"] + 12 * ncon
    lay["cbA"] = lay["T"] + 12 * ncon + (0 if wrench_alias else wrench_floats)
    lay["aforce"] = lay["x"]            # actuator forces live only inside smooth_forces()
    lay["tmpv"] = lay["qacc_smooth"]    # solve() temp (g_k): qacc_smooth is consumed (into registers) before the first CG solve
    for k, v in lay.items():
        S("o_" + k, v)
    # the observation row is staged over crb/LD/T at the end of the step, shifted by up to 3 floats to the 16-byte phase of its
    # destination row (csrc/bt_programs.h::bt_write_obs)
    off = max(off, lay["crb"] + obs_size + 3)
    S("smem_floats", off + (-off) % 4)
    S("sync_mode", default_sync_mode(int(t["smem_floats"][0])))
    # ------------------------------------------------------------------ constant records shared by the warps of a CTA
    # The constant records that the lane-parallel passes of EVERY substep reach through a chain of dependent loads (body_frame:
    # body_rec -> jnt_rec -> qpos; smooth_forces: act_rec -> wrap_rec, dof_rec -> dofact_rec) are one contiguous table, `sh_tab`; the
    # kernels copy its first `sh_stage_floats` floats into the shared memory that the per-environment slices leave free and read them
    # from there (BtEnv::crec).  Why: with 14 x 16 KB of scratch the L1 is 22 KB against ~31 KB of tables touched per substep -- a
    # cyclic pattern that an LRU cache misses every time (ncu, rodent: 4.2 M L1 miss sectors per launch at 14 warps per CTA against
    # 1.8 M at 12, where the carve-out leaves 55 KB of L1) -- and the warps of a CTA all wait for the same first touch.
    # Tables are taken greedily in the priority order below while they fit (they then form the prefix of sh_tab); the rest is read
    # from global memory as before.  Measured, same-box A/B (profiles/r2l_staged_records.txt): rodent body_rec + bl_rec +1.1 %,
    # body_rec + jnt_rec instead +0.65 %, + wrap_rec + dofact_rec +0.9 %; first-level records with independent loads (bl_rec,
    # act_rec, dof_rec) gain nothing.
    priority = ("body_rec", "jnt_rec", "wrap_rec", "dofact_rec", "bl_rec")
    env_bytes = 4 * int(t["smem_floats"][0])
    need_ds, need_cs = (nv + 31) // 32, max((ncon + 31) // 32, 1)
    max_warps = MAX_WARPS_PER_CTA if (need_ds <= 3 and need_cs <= 1) else 8      # csrc/bt_ops.h: BT_VARIANT_MAX_WARPS
    budget = SMEM_BYTES_PER_SM - min(max_warps, SMEM_BYTES_PER_SM // env_bytes) * env_bytes   # what the environments leave free
    # (the 2-slot kernel variant -- nv <= 64: the fly models -- reads these records from global memory: csrc/bt_impl.h, kSmallModel)
    nstage = 0 if need_ds <= 2 else int(os.environ.get("BT_STAGE", len(priority)))   # BT_STAGE = n: only the first n candidates
    padded = lambda k: t[k].size + (-t[k].size) % 4
    staged, used = [], 0
    for k in priority[:nstage]:
        if 4 * (used + padded(k)) <= budget:
            staged.append(k); used += padded(k)
    parts, so = [], 0
    for k in staged + [k for k in priority if k not in staged]:
        v = np.asarray(t[k], dtype=np.float32)
        S("sho_" + k, so)
        parts.append(np.concatenate([v, np.zeros(padded(k) - v.size, np.float32)]))
        so += padded(k)
    t["sh_tab"] = np.concatenate(parts)
    stage = used
    S("sh_stage_floats", stage)
    return t


SMEM_BYTES_PER_SM = 227 * 1024   # opt-in dynamic shared memory of one sm_100a CTA
MAX_WARPS_PER_CTA = 16           # csrc/bt_ops.h: BT_MAX_WARPS


def default_sync_mode(smem_floats: int) -> int:
    """Phase alignment of the warps of a CTA (bits, csrc/bt_impl.h::substep).  The warps own different environments but must walk
    the SAME code at the same time (6 KB L0 / 32 KB L1.5 instruction caches against 240 KB of SASS): 1 = barrier over all warps of
    the CTA at the start of every substep; 32 / 64 / 128 = only among the warps of a contiguous half / of equal parity / of equal
    (index mod 4); 2, 4, 8, 16 = further alignment points after the tree pass / before each factorisation / before collision /
    in every CG pass, CTA-wide or -- with 1024 -- among the warps of equal parity; 2048 / 4096 keep only the first / the second of
    the two `4` points; 0 = none.  BT_SYNC overrides.

    Measured on one B200, 8192 envs (env-steps/s; profiles/r2l_alignment_points.txt, r2l_split_alignment.txt):
      rodent        0: 1.39 M (r1q)   1: 2.756 M   32: 2.777 M   64: 2.781 M   128: 2.61 M (r1af)
                    64: 3.075 M   64+1024+4: 3.116 M   64+1024+4+4096: 3.115 M   64+1024+4+2048: 3.075 M   64+1024+16: 3.092 M
      fly, free     64: 3.34 M    64+1024+4+4096: 3.72 M  (any one extra point gives the same)
      fly, tethered 64: 5.82 M    64+1024+4+4096: 6.15 M   64+1024+16: 6.20 M
      two rodents (8 warps per CTA)  64: 778 k   64+1024+4+4096: 768 k   64+1024+16: 772 k
    i.e. the environments of a group leave the CG loop at different times (iteration and line-search counts differ) and one
    re-alignment before the Euler factorisation pays when a group has 7-8 warps, not when it has 4."""
    env = os.environ.get("BT_SYNC")
    if env is not None:
        mode = int(env)
        if bin(mode & (32 | 64 | 128)).count("1") > 1:
            # two different groupings would meet on the same named barriers with different arrival counts: a deadlock
            raise ValueError("BT_SYNC: at most one of the group-alignment bits 32 / 64 / 128 may be set")
        if (mode & 1024) and (mode & (32 | 128)):
            raise ValueError("BT_SYNC: bit 1024 aligns the warps of equal parity and goes with bit 64 (or 1) only")
        return mode
    envs_per_sm = min(MAX_WARPS_PER_CTA, SMEM_BYTES_PER_SM // (4 * smem_floats))
    return (64 | 1024 | 4 | 4096) if envs_per_sm >= 12 else 64


def model_dims(t: Dict[str, np.ndarray]) -> dict:
    g = lambda k: int(t[k][0])
    return dict(nq=g("nq"), nv=g("nv"), nu=g("nu"), na=g("na"), nbody=g("nbody"), ncon=g("ncon"),
                obs_size=g("obs_size"), smem_floats=g("smem_floats"), n_frames=g("n_frames"), cone=g("cone"))

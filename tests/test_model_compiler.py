"""Pins the in-repo MJCF mini-compiler (brax_tracking_b200/mjcf.py): oracle and product read the SAME compiled tables, so a
compiler error is invisible to every parity test.  Pins: the model facts of SURVEY.md Appendix C (derived
there by reading the MJCF files), the total mass re-derived from the XML by an independent walk written here (its own default-class
resolution and volume formulas; nothing imported from mjcf.py), left / right symmetry of the compiled inertias, and the 0.9
rescale (rodent.py:60-64) multiplying every mass by 0.9^3.  The committed assets (`assets/*.npz`) are checked against a fresh
compile when the reference checkout is present, and against the Appendix C facts always."""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from brax_tracking_b200 import assets, mjcf, model

REF = os.environ.get("BT_REFERENCE_ROOT", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "assets", "rodent.xml")), reason="reference checkout not present")


# ---------------------------------------------------------------------------------- Appendix C facts on the committed assets
def _tree_facts(m):
    a = m.a
    depth = int(a["body_depth"].max())
    ddepth = int(a["dof_depth"].max()) + 1
    carry = len(set(int(b) for b in a["dof_bodyid"]))
    gtype = a["geom_type"]
    col = sorted({int(g) for p, n in zip(a["pair_geom"], a["pair_ncon"]) for g in p if gtype[g] != mjcf.GEOM_PLANE})
    leaf_chains = sum(1 for i in range(m.nv) if not any(a["dof_parentid"][j] == i for j in range(m.nv)))
    return dict(depth=depth, dof_chain_depth=ddepth, bodies_with_dofs=carry, collidable=col, leaf_chains=leaf_chains)


def test_rodent_matches_survey_appendix_c1():
    m = assets.load_model("rodent")
    a = m.a
    assert (m.nbody, m.njnt, m.nq, m.nv, m.nu, m.na, m.ntendon, m.ngeom) == (67, 68, 74, 73, 38, 38, 8, 101)
    f = _tree_facts(m)
    assert f["depth"] == 39 and f["dof_chain_depth"] == 36 and f["bodies_with_dofs"] == 52 and f["leaf_chains"] == 6
    assert m.nM == 1119                                                        # tree-sparse qM non-zeros (dense: 5329)
    col = f["collidable"]
    assert len(col) == 16 and sum(a["geom_type"][g] == mjcf.GEOM_CAPSULE for g in col) == 14 \
        and sum(a["geom_type"][g] == mjcf.GEOM_ELLIPSOID for g in col) == 2
    assert int((a["geom_type"] == mjcf.GEOM_PLANE).sum()) == 1
    assert int(a["pair_ncon"].sum()) == 30 and set(a["pair_condim"]) == {3}    # 14 x 2 + 2 x 1 floor contacts
    nlimit = int(((a["jnt_limited"] != 0) & (a["jnt_type"] == mjcf.JNT_HINGE)).sum())
    assert nlimit == 67 and nlimit + 4 * 30 == 187                             # nefc: limits + pyramidal rows
    assert int((a["actuator_trntype"] == mjcf.TRN_TENDON).sum()) == 8
    np.testing.assert_allclose(a["pair_friction"][:, 0], 1.5)                  # paw priority 1 wins (rodent.xml:21-28)
    np.testing.assert_allclose(a["pair_solref"], [[0.005, 1.0]] * len(a["pair_ncon"]))
    assert m.timestep == 0.002 and m.cone == mjcf.CONE_PYRAMIDAL and (m.iterations, m.ls_iterations) == (4, 4)


def test_fly_matches_survey_appendix_c2():
    ff, ft = assets.load_model("fly_free"), assets.load_model("fly_tethered")
    assert (ff.nbody, ff.njnt, ff.nq, ff.nv, ff.nu, ff.na, ff.nM) == (68, 37, 43, 42, 36, 0, 363)
    assert (ft.nq, ft.nv, ft.nu, ft.nM) == (36, 36, 36, 126) and ft.nbody == 68
    assert int(ff.a["dof_depth"].max()) + 1 == 12
    # tethered: six independent 6-dof chains (block-diagonal qM)
    roots = [i for i in range(ft.nv) if ft.a["dof_parentid"][i] < 0]
    assert len(roots) == 6 and all(int(ft.a["dof_subtreenum"][i]) == 6 for i in roots)
    for m in (ff, ft):
        a = m.a
        assert m.cone == mjcf.CONE_ELLIPTIC and int(a["pair_ncon"].sum()) == 27
        dims = sorted(int(d) for d, n in zip(a["pair_condim"], a["pair_ncon"]) for _ in range(n))
        assert dims == [1] * 15 + [3] * 12                                      # 15 claw-claw (condim 1) + 6 x 2 floor (condim 3)
        np.testing.assert_allclose(a["pair_margin"] - a["pair_gap"], 0.0, atol=1e-12)   # margin = gap = 5e-4
        assert abs(m.density - 0.00128) < 1e-12 and abs(m.viscosity - 0.000185) < 1e-12


def test_pair_matches_survey_appendix_c3():
    m = assets.load_model("rodent_pair")
    assert (m.nbody, m.nq, m.nv, m.nu, m.na, m.ntendon, m.nM) == (133, 148, 146, 60, 60, 0, 2238)
    a = m.a
    plane = a["geom_type"][a["pair_geom"][:, 0]] == mjcf.GEOM_PLANE
    assert int(a["pair_ncon"][plane].sum()) == 114 and int(a["pair_ncon"][~plane].sum()) == 12     # 2 (27 x 2 + 3) + opened pairs
    roots = [i for i in range(m.nv) if a["dof_parentid"][i] < 0]
    assert roots == [0, 73]                                                     # block-diagonal 2 x 73
    for g1, g2 in a["pair_geom"][~plane]:                                       # every opened pair joins the two animals
        assert {int(a["body_rootid"][a["geom_bodyid"][g1]]), int(a["body_rootid"][a["geom_bodyid"][g2]])} == {1, 67}
    from brax_tracking_b200 import configs, presets
    t = model.pack(m, configs.resolve(m, presets.ENV_ARGS["rodent_pair"]), presets.load("rodent_pair")[2])
    assert int(t["obs_size"][0]) == 148 + 146 + 2 * (15 + 20 + 165 + 270) == 1234


# ---------------------------------------------------------------------------------- independent mass walk over the XML
def _defaults(root):
    """class name -> attribute dict for <geom>, inherited along the nesting of <default> elements."""
    out = {}

    def walk(d, inherited):
        mine = dict(inherited)
        g = d.find("geom")
        if g is not None:
            mine.update(g.attrib)
        out[d.get("class", "main")] = mine
        for c in d.findall("default"):
            walk(c, mine)
    for d in root.findall("default"):
        walk(d, {})
    return out


def _volume(at):
    size = [float(x) for x in at.get("size", "0").split()]
    t = at.get("type", "sphere")
    if "fromto" in at:
        ft = [float(x) for x in at["fromto"].split()]
        half = 0.5 * math.dist(ft[:3], ft[3:])
        size = [size[0], half]
    if t == "sphere":
        return 4 / 3 * math.pi * size[0] ** 3
    if t == "capsule":
        return math.pi * size[0] ** 2 * 2 * size[1] + 4 / 3 * math.pi * size[0] ** 3
    if t == "ellipsoid":
        return 4 / 3 * math.pi * size[0] * size[1] * size[2]
    if t == "box":
        return 8 * size[0] * size[1] * size[2]
    if t == "cylinder":
        return math.pi * size[0] ** 2 * 2 * size[1]
    if t == "plane":
        return 0.0
    raise NotImplementedError(t)


def _xml_total_mass(path):
    root = ET.parse(path).getroot()
    dflt = _defaults(root)
    total = 0.0

    def walk(elem, childclass):
        nonlocal total
        for c in elem:
            if c.tag == "geom":
                at = dict(dflt.get("main", {}))
                at.update(dflt.get(c.get("class", childclass) or "main", {}))
                at.update(c.attrib)
                if elem.tag != "worldbody":
                    total += float(at["mass"]) if "mass" in at else float(at.get("density", 1000.0)) * _volume(at)
            elif c.tag == "body":
                walk(c, c.get("childclass", childclass))
    wb = root.find("worldbody")
    walk(wb, wb.get("childclass"))
    return total


@needs_ref
def test_total_mass_matches_an_independent_walk_over_the_xml():
    path = os.path.join(REF, "assets", "rodent.xml")
    want = _xml_total_mass(path)
    m = mjcf.compile_mjcf(path, overrides=dict(iterations=4, ls_iterations=4))
    assert abs(m.body_mass.sum() - want) < 1e-9 * want, (m.body_mass.sum(), want)
    assert 0.2 < want < 0.3                                                     # a ~230 g rat
    scaled = mjcf.compile_mjcf(path, scale_factor=0.9, overrides=dict(iterations=4, ls_iterations=4))
    np.testing.assert_allclose(scaled.body_mass, 0.729 * m.body_mass, rtol=1e-9)                 # rescale_subtree(0.9): mass x 0.9^3
    np.testing.assert_allclose(scaled.body_inertia, 0.9 ** 5 * m.body_inertia, rtol=1e-9)       # ... inertia x 0.9^5
    np.testing.assert_allclose(scaled.body_pos, 0.9 * m.body_pos, rtol=1e-12, atol=1e-15)
    committed = assets.load_model("rodent")                                       # the committed asset IS this compile
    for k in ("body_mass", "body_inertia", "body_pos", "body_quat", "jnt_axis", "jnt_range", "dof_invweight0", "body_invweight0"):
        np.testing.assert_array_equal(committed.a[k], scaled.a[k], err_msg=k)


def test_left_right_symmetry_of_the_compiled_rodent():
    m = assets.load_model("rodent")
    names = m.names["body"]
    pairs = [(names.index(n), names.index(n[:-2] + "_R")) for n in names if n.endswith("_L") and n[:-2] + "_R" in names]
    assert len(pairs) == 9
    for l, r in pairs:
        assert abs(m.body_mass[l] - m.body_mass[r]) < 1e-6 * m.body_mass[l], names[l]
        np.testing.assert_allclose(m.body_inertia[l], m.body_inertia[r], rtol=1e-4, err_msg=names[l])
        # mirrored in y (the XML itself places a few mirrored geoms a fraction of a millimetre apart, e.g. lower_leg_L/R_collision)
        np.testing.assert_allclose(m.body_ipos[l] * [1, -1, 1], m.body_ipos[r], atol=5e-4, err_msg=names[l])
    # the constants MuJoCo derives at qpos0 (mj_setConst): dof_invweight0 = diag(M^-1), equal on mirrored joints
    jn = m.names["joint"]
    for n in jn:
        if n.endswith("_L") and n[:-2] + "_R" in jn:
            dl, dr = m.jnt_dofadr[jn.index(n)], m.jnt_dofadr[jn.index(n[:-2] + "_R")]
            assert abs(m.dof_invweight0[dl] - m.dof_invweight0[dr]) < 1e-2 * m.dof_invweight0[dl], n   # (asset asymmetry: 0.2 %)
    # ... and it IS diag(M^-1) at qpos0 (mj_setConst), with M from a different construction in a different language: the C oracle's
    # composite-rigid-body recursion (the compiler itself sums body Jacobians in numpy)
    import oracle as oracle_mod
    o = oracle_mod.Oracle(m, np.float64)
    o.set_state(m.qpos0, np.zeros(m.nv))
    o.forward()
    M = np.array(o.d.qM).reshape(m.nv, m.nv)
    M = np.tril(M) + np.tril(M, -1).T if not np.allclose(M, M.T) else M
    assert np.linalg.eigvalsh(M).min() > 0
    np.testing.assert_allclose(np.diag(np.linalg.inv(M))[6:], m.dof_invweight0[6:], rtol=1e-6)
    assert abs(m.meaninertia - np.trace(M) / m.nv) < 1e-9 * m.meaninertia

"""Structural invariants of the scheduling tables model.py::pack hands to the kernels (chain descriptors, sweep passes,
contact segments, wrench groups, pose-composition work lists).  Pure host logic: no oracle, no GPU."""
import numpy as np
import pytest

import common

MODELS = ["rodent", "fly_free", "fly_tethered", "rodent_pair"]


def _desc(row):
    return dict(k0=row[0], kb=row[1], pc=row[2], c=row[3], cadr=row[4], nch=row[5] & 0xffff, nseg=row[5] >> 16, sadr=row[6], pdof=row[7])


@pytest.mark.parametrize("name", MODELS)
def test_chain_descriptors_cover_the_dof_tree(name):
    m, cfg, clip, t = common.setup(name)
    nv = m.nv
    desc = t["chain_desc"].reshape(-1, 8)
    par = np.asarray(m.a["dof_parentid"])
    seen = np.zeros(nv, int)
    for c, row in enumerate(desc):
        d = _desc(row)
        assert d["c"] == c and d["pdof"] == par[d["k0"]]
        seen[d["k0"]:d["kb"] + 1] += 1
        for k in range(d["k0"] + 1, d["kb"] + 1):           # a chain is a single-child path of consecutive dofs
            assert par[k] == k - 1
        kids = t["cchild_id"][d["cadr"]:d["cadr"] + d["nch"]]
        for ch in kids:                                       # child chains hang off the END of their parent chain
            assert desc[ch][2] == c and par[desc[ch][0]] == d["kb"]
        ends = t["seg_end"][d["sadr"]:d["sadr"] + d["nseg"]]  # segments tile the chain; every inner end is a contact dof
        assert list(ends) == sorted(ends) and ends[-1] == d["kb"] and ends[0] >= d["k0"]
        cbs = t["seg_cb"][d["sadr"]:d["sadr"] + d["nseg"]]
        assert all(cb >= 0 for cb in cbs[:-1])
        for e, cb in zip(ends, cbs):
            if cb >= 0:
                assert t["cb_lastdof"][cb] == e
    assert (seen == 1).all()


@pytest.mark.parametrize("name", MODELS)
def test_sweep_passes_respect_the_dependencies(name):
    m, cfg, clip, t = common.setup(name)
    desc = t["chain_desc"].reshape(-1, 8)
    for key, width in (("hpass_desc", 32), ("apass_desc", 4)):
        rows = t[key].reshape(-1, width, 8)
        pass_of = {}
        for p, row in enumerate(rows):
            for lane in row:
                if lane[1] >= lane[0]:
                    assert np.array_equal(lane, desc[lane[3]])
                    pass_of[int(lane[3])] = p
        assert sorted(pass_of) == list(range(len(desc)))      # every chain exactly once
        for c, p in pass_of.items():
            pc = desc[c][2]
            if pc >= 0:                                       # hpass: root-most first; apass: deepest first
                assert pass_of[pc] < p if key == "hpass_desc" else pass_of[pc] > p


@pytest.mark.parametrize("name", MODELS)
def test_wrench_groups_and_merge_lists(name):
    m, cfg, clip, t = common.setup(name)
    par = np.asarray(m.a["dof_parentid"])
    ncb = int(t["ncb"][0])
    below = [set() for _ in range(m.nv)]                      # contact bodies whose ancestor chain contains the dof
    for cb in range(ncb):
        d = int(t["cb_lastdof"][cb])
        while d >= 0:
            below[d].add(cb); d = par[d]
    for d in range(m.nv):
        g = t["dof_wgrp"][d]
        got = set(t["wgrp_cb"][t["wgrp_adr"][g]:t["wgrp_adr"][g + 1]]) if g >= 0 else set()
        assert got == below[d]
        if g >= 0 and int(t["wgrp_contig"][0]):   # the packed (first | count << 16) form the kernel reads
            rg = int(t["wgrp_rng"][g])
            assert set(range(rg & 0xffff, (rg & 0xffff) + (rg >> 16))) == below[d]
    assert int(t["wgrp_contig"][0]) == 1           # DFS numbering of the contact bodies: true for every shipped model
    # link records: every moving body is either the record of its dof or merged into it exactly once
    lastdof = np.asarray(m.a["body_lastdof"])
    owner = {}
    for d in range(m.nv):
        if t["dof_irec"][d] >= 0:
            owner[int(t["dof_irec"][d])] = d
    merged = {}
    for r in range(int(t["nmerge"][0])):
        for s in t["merge_src"][t["merge_adr"][r]:t["merge_adr"][r + 1]]:
            merged[int(s)] = int(t["merge_dst"][r])
    for b in range(1, m.nbody):
        if lastdof[b] >= 0:
            assert (b in owner and owner[b] == lastdof[b]) ^ (b in merged and owner[merged[b]] == lastdof[b])


@pytest.mark.parametrize("name", MODELS)
def test_pose_composition_work_lists(name):
    """Replays the pointer-jumping rounds symbolically: every body ends at the world frame in buffer 0 (xpos / xquat)."""
    m, cfg, clip, t = common.setup(name)
    nb, R = m.nbody, int(t["nbanc"][0])
    parent = np.asarray(m.a["body_parentid"])
    depth = np.zeros(nb, int)
    for b in range(1, nb):
        depth[b] = depth[parent[b]] + 1
    # buf[which][b] = how many levels above b its stored pose is expressed in (depth[b] = world); -1 = stale
    buf = [np.full(nb, -1), np.full(nb, -1)]
    buf[R & 1][1:] = 1                                        # body_frame: pose relative to the parent
    buf[R & 1][depth == 1] = 1
    for r in range(R):
        src, dst = buf[(R - r) & 1], buf[(R - 1 - r) & 1]
        new = dst.copy()
        for w in t["cmp_item"][t["cmp_adr"][r]:t["cmp_adr"][r + 1]]:
            b, a = w & 0xfff, (w >> 12) & 0xfff
            assert src[b] > 0
            if w & (1 << 29):
                assert src[b] == depth[b]                     # copy-forward of a finished pose
                new[b] = src[b]
            else:
                assert src[a] > 0 and depth[a] == depth[b] - src[b]   # a is the body the stored pose is relative to
                new[b] = src[b] + src[a]
                assert bool(w & (1 << 28)) == (new[b] == depth[b])
        dst[:] = new
    assert (buf[0][1:] == depth[1:]).all()


@pytest.mark.parametrize("name", MODELS)
def test_regions_with_128_bit_accesses_are_aligned(name):
    m, cfg, clip, t = common.setup(name)
    for k in ("o_cdof", "o_xquat", "o_T", "o_pvec"):          # dof records, quaternions (both pose buffers), cvel|cacc records
        assert int(t[k][0]) % 4 == 0, k
    assert int(t["smem_floats"][0]) % 4 == 0                   # every environment's block starts 16-byte aligned
    need = ((3 * m.nbody + 3) & ~3) + 4 * m.nbody              # second pose buffer inside T
    views = ("o_cbA", "o_wrench")                               # sub-views of T / of Dd | cbJ, not regions of their own
    nxt = min(int(v[0]) for k, v in t.items() if k.startswith("o_") and k not in views and int(v[0]) > int(t["o_T"][0]))
    assert int(t["o_T"][0]) + need <= nxt


def test_default_alignment_mode_follows_the_environments_per_sm(monkeypatch):
    """model.default_sync_mode: one extra parity-group barrier before the Euler factorisation when a barrier group has 7-8 warps
    (rodent 14, flies 16 environments per SM), the substep barrier alone for the two-rodent model (6 per SM); BT_SYNC overrides and
    rejects groupings that would meet on the same named barriers with different arrival counts."""
    from brax_tracking_b200 import model
    monkeypatch.delenv("BT_SYNC", raising=False)
    for name, want in (("rodent", 64 | 1024 | 4 | 4096), ("fly_free", 64 | 1024 | 4 | 4096), ("fly_tethered", 64 | 1024 | 4 | 4096),
                       ("rodent_pair", 64)):
        t = common.setup(name)[3]
        assert int(t["sync_mode"][0]) == want, name
        assert model.default_sync_mode(int(t["smem_floats"][0])) == want
    monkeypatch.setenv("BT_SYNC", "1")
    assert model.default_sync_mode(4000) == 1
    for bad in ("96", "192", str(32 | 1024 | 4), str(128 | 1024 | 16)):
        monkeypatch.setenv("BT_SYNC", bad)
        with pytest.raises(ValueError):
            model.default_sync_mode(4000)


@pytest.mark.parametrize("name", MODELS)
def test_staged_record_table_and_factor_pass_schedule(name):
    """sh_tab holds body_rec, jnt_rec, wrap_rec, dofact_rec, bl_rec (each padded to 16 bytes), the staged ones first: taken
    greedily in that priority order while they fit what the environments leave free of the 227 KB (staging never costs an
    environment); nothing is staged for the 2-slot (fly) kernel variant.  The factor-sweep passes (critical-path list scheduling)
    are never longer than the schedule by depth level."""
    from brax_tracking_b200 import model
    m, cfg, clip, t = common.setup(name)
    prio = ("body_rec", "jnt_rec", "wrap_rec", "dofact_rec", "bl_rec")
    pad = lambda k: t[k].size + (-t[k].size) % 4
    offs = {k: int(t["sho_" + k][0]) for k in prio}
    for k in prio:
        assert offs[k] % 4 == 0 and np.array_equal(t["sh_tab"][offs[k]:offs[k] + t[k].size], t[k])
    assert t["sh_tab"].size == sum(pad(k) for k in prio)
    assert sorted((offs[k], offs[k] + pad(k)) for k in prio) == [(a, b) for a, b in zip(np.cumsum([0] + [pad(k) for k in sorted(prio, key=offs.get)][:-1]),
                                                                                       np.cumsum([pad(k) for k in sorted(prio, key=offs.get)]))]
    stage = int(t["sh_stage_floats"][0])
    env_bytes = 4 * int(t["smem_floats"][0])
    small = (m.nv + 31) // 32 <= 2
    max_warps = 16 if ((m.nv + 31) // 32 <= 3 and (int(t["ncon"][0]) + 31) // 32 <= 1) else 8
    envs = min(max_warps, model.SMEM_BYTES_PER_SM // env_bytes)
    budget = model.SMEM_BYTES_PER_SM - envs * env_bytes
    want, used = [], 0
    for k in ([] if small else prio):
        if 4 * (used + pad(k)) <= budget:
            want.append(k); used += pad(k)
    assert stage == used and 4 * stage + envs * env_bytes <= model.SMEM_BYTES_PER_SM
    assert {k for k in prio if offs[k] < stage} == set(want)                  # the staged tables are exactly the prefix
    if name == "rodent":
        assert want == ["body_rec", "jnt_rec", "wrap_rec", "dofact_rec"]
    # factor passes: serial dof steps = sum over passes of the longest chain; by depth level for comparison
    desc = t["chain_desc"].reshape(-1, 8)
    ap = t["apass_desc"].reshape(-1, 4, 8)
    steps = sum(max(int(r[1] - r[0] + 1) for r in rows) for rows in ap)
    depth = {}
    for c in range(len(desc)):
        pc, d = int(desc[c][2]), 0
        while pc >= 0:
            d += 1; pc = int(desc[pc][2])
        depth[c] = d
    by_level = 0
    for L in sorted(set(depth.values())):
        cs = [c for c in range(len(desc)) if depth[c] == L]
        for q in range(0, len(cs), 4):
            by_level += max(int(desc[c][1] - desc[c][0] + 1) for c in cs[q:q + 4])
    assert steps <= by_level
    if name == "rodent":
        assert (steps, by_level) == (38, 39)
    if name == "rodent_pair":
        assert (steps, by_level) == (53, 72)

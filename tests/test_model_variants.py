"""Model-changing task kwargs of the reference (tasks/base.py:96-107, shadow_hand.py:162-182,272-311):
reduced_action_space, attachment_yaw, forearm_dofs subsets.  CPU part: the flattened scenes, the oracle on them (its
mass matrix against the independent numpy restatement, bodies without joints included), the C-ABI model packer
accepting them (host only, no GPU)."""
import numpy as np
import pytest

import helpers
import np_ref
from oracle import oracle as O
from robopianist_b200 import native
from robopianist_b200.scene import cache, compile as C
from robopianist_b200.scene import mathutil as mu
from robopianist_b200.suite import env as E

VARIANTS = {
    "ras": dict(reduced_action_space=True),
    "yaw10": dict(attachment_yaw=10.0),
    "fdx": dict(forearm_dofs=("forearm_tx",)),
    "ras_fd": dict(forearm_dofs=(), reduced_action_space=True),
    "hc": dict(disable_hand_collisions=False),
    "hc_ras": dict(disable_hand_collisions=False, reduced_action_space=True),
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_scene_layout(name):
    kw = VARIANTS[name]
    fm = cache.load_scene(gravity_compensation=True, **kw)
    full = cache.load_scene(gravity_compensation=True)
    ras = bool(kw.get("reduced_action_space"))
    nfd = len(kw.get("forearm_dofs", ("forearm_tx", "forearm_ty")))
    nl = 24 - (3 if ras else 0) + nfd
    assert fm.meta["nv"] == 88 + 2 * nl and fm.meta["nu"] == 2 * (20 - (3 if ras else 0) + nfd)
    assert fm.arrays["hand_obs_dofid"].shape == (2, nl)
    names = fm.meta["joint_names"]
    if ras:  # shadow_hand.py:73-79,162-179
        for side in ("rh", "lh"):
            for dof in C.REDUCED_ACTION_SPACE_EXCLUDED_DOFS:
                assert f"{side}_shadow_hand/{side}_{dof}" not in names
                assert f"{side}_shadow_hand/{side}_A_{dof}" not in fm.meta["actuator_names"]
            j = names.index(f"{side}_shadow_hand/{side}_THJ2")
            assert np.allclose(fm.jnt_range[j], C.REDUCED_THUMB_RANGE)
            a = fm.meta["actuator_names"].index(f"{side}_shadow_hand/{side}_A_THJ2")
            assert np.allclose(fm.actuator_ctrlrange[a], C.REDUCED_THUMB_RANGE)
        # the bodies of the removed joints stay, without joints
        jointless = [n for n, k in zip(fm.meta["body_names"], fm.body_dofnum) if k == 0 and "shadow_hand" in n]
        assert sum(n.endswith(("lfmetacarpal", "thbase", "thdistal")) for n in jointless) == 6
        assert fm.meta["nbody"] == full.meta["nbody"] and fm.meta["ngeom"] == full.meta["ngeom"]
    if "attachment_yaw" in kw:  # base.py:175-183: rotate_by (about z, +yaw right / -yaw left) times the hand quaternion
        for side, sign in (("rh", 1.0), ("lh", -1.0)):
            b = fm.meta["body_names"].index(f"{side}_shadow_hand/{side}_forearm")
            ang = np.radians(sign * kw["attachment_yaw"])
            want = mu.quat_normalize(mu.quat_mul(np.array([np.cos(ang / 2), 0, 0, np.sin(ang / 2)]), np.array(C.HAND_QUAT, float)))
            assert np.allclose(fm.body_quat[b], want, atol=1e-15)
            assert not np.allclose(fm.body_quat[b], full.body_quat[b])
    # observation joint order: XML joints first, forearm dofs last (shadow_hand.py:159-183)
    for h, side in enumerate(("rh", "lh")):
        obs_names = [names[j].split("/")[1] for j in fm.arrays["hand_obs_dofid"][h]]
        assert obs_names[len(obs_names) - nfd:] == list(kw.get("forearm_dofs", ("forearm_tx", "forearm_ty")))
        assert all(n.startswith(side + "_") for n in obs_names[: len(obs_names) - nfd])


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_mass_matrix_on_variants(name):
    """Bodies without joints (reduced action space, a forearm without dofs) go through the oracle's general body
    tree; its composite-inertia mass matrix has to equal the Jacobian-sum restatement."""
    fm = cache.load_scene(gravity_compensation=True, **VARIANTS[name])
    om = O.OracleModel(fm)
    d = O.OracleData(om)
    rng = np.random.default_rng(3)
    nv = fm.meta["nv"]
    q = fm.jnt_range[:, 0] + rng.random(nv) * (fm.jnt_range[:, 1] - fm.jnt_range[:, 0])
    d.qpos[:] = q
    for fn in ("orc_kinematics", "orc_crb"):
        d.call(fn)
    assert np.abs(np_ref.mass_matrix(fm, q) - d.M).max() < 1e-12
    # and it steps (random controls, a few control steps) without leaving the finite range
    d2 = O.OracleData(om); d2.reset()
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    for _ in range(5):
        d2.control_step(lo + (hi - lo) * rng.random(fm.meta["nu"]))
    assert np.isfinite(d2.qpos).all() and np.isfinite(d2.qvel).all()


@pytest.mark.parametrize("name", list(VARIANTS))
def test_packer_accepts_variants(name):
    """rp_model_create (host side, no GPU): variable dof count per hand, jointless bodies fused into their parents."""
    fm = cache.load_scene(gravity_compensation=True, **VARIANTS[name])
    goal, fing = helpers.song_tables()
    model = native.Model(fm, goal, fing, kind=0, n_steps_lookahead=1)
    nl = fm.arrays["hand_obs_dofid"].shape[1]
    assert model.act_dim == fm.meta["nu"] + 1
    assert model.obs_dim == 88 + 1 + 2 * nl + 2 * 89 + 10
    lay = E.obs_layout(True, 1, False, nl)
    assert lay["rh_shadow_hand/joints_pos"] == slice(89, 89 + nl) and next(reversed(lay.values())).stop == model.obs_dim


def test_unsupported_kwargs_still_raise():
    kw = dict(E._TASK_KWARGS, primitive_fingertip_collisions=True, disable_hand_collisions=True)
    E._unsupported(dict(kw, reduced_action_space=True, attachment_yaw=5.0, randomize_hand_positions=True))  # accepted now
    with pytest.raises(NotImplementedError):
        E._unsupported(dict(kw, primitive_fingertip_collisions=False))
    E._unsupported(dict(kw, disable_hand_collisions=False))  # accepted now (sphere / capsule pairs collide)
    with pytest.raises(NotImplementedError):
        E._unsupported(dict(kw, forearm_dofs=("forearm_roll",)))


def test_compile_matches_cached_variants():
    """The committed variant scenes are what compile.py produces from the assets (build container only)."""
    if not C.DEFAULT_ASSET_DIR.exists():
        pytest.skip("Shadow Hand assets not present")
    for name, kw in VARIANTS.items():
        fresh = C.compile_scene(gravity_compensation=True, **kw)
        cached = cache.load_scene(gravity_compensation=True, **kw)
        for k, v in fresh.arrays.items():
            assert np.array_equal(v, cached.arrays[k]), (name, k)


def test_hand_collision_scene_and_pair_filters():
    """disable_hand_collisions=False (the reference default; piano_with_shadow_hands.py:476-489 is what switches it off):
    hand geoms keep contype = conaffinity = 1; the oracle's pair list obeys mj_collision's body-pair filters."""
    import ctypes as Ct
    fm = cache.load_scene(gravity_compensation=True, disable_hand_collisions=False)
    off = cache.load_scene(gravity_compensation=True)
    hand = fm.geom_hand >= 0
    assert (fm.geom_contype[hand] == 1).all() and (fm.geom_conaffinity[hand] == 1).all()
    assert (off.geom_contype[hand] == 1).all() and (off.geom_conaffinity[hand] == 0).all()
    names = fm.meta["body_names"]
    ex = {tuple(sorted((names[a].split("/")[1], names[b].split("/")[1]))) for a, b in fm.arrays["exclude_bodies"]}
    assert ex == {("rh_forearm", "rh_wrist"), ("rh_thmiddle", "rh_thproximal"), ("lh_forearm", "lh_wrist"), ("lh_thmiddle", "lh_thproximal")}
    om = O.OracleModel(fm)
    L = O.lib()
    L.orc_npair.argtypes = [Ct.c_void_p]; L.orc_pair_get.argtypes = [Ct.c_void_p, Ct.c_int, Ct.POINTER(Ct.c_int), Ct.POINTER(Ct.c_int)]
    n = L.orc_npair(om.ptr)
    pairs = []
    for i in range(n):
        a, b = Ct.c_int(), Ct.c_int()
        L.orc_pair_get(om.ptr, i, Ct.byref(a), Ct.byref(b))
        pairs.append((a.value, b.value))
    assert pairs == sorted(pairs) and len(set(pairs)) == n
    cov = [g for g in range(fm.meta["ngeom"]) if fm.geom_hand[g] >= 0 and fm.geom_collide[g]]
    assert len(cov) == 36
    parent = fm.body_parentid
    want = 0
    for i, g1 in enumerate(cov):
        for g2 in cov[i + 1:]:
            b1, b2 = fm.geom_bodyid[g1], fm.geom_bodyid[g2]
            ok = b1 != b2 and parent[b1] != b2 and parent[b2] != b1  # every hand body of this model has a joint: weld == body
            ok = ok and tuple(sorted((names[b1].split("/")[1], names[b2].split("/")[1]))) not in ex
            want += ok
            assert ((g1, g2) in pairs) == ok, (fm.meta["geom_names"][g1], fm.meta["geom_names"][g2])
    assert n == want == 596
    assert L.orc_npair(O.OracleModel(off).ptr) == 0
    # with the reduced action space three bodies per hand are welded to their parents: more filtered pairs
    ras = cache.load_scene(gravity_compensation=True, disable_hand_collisions=False, reduced_action_space=True)
    n_ras = L.orc_npair(O.OracleModel(ras).ptr)
    assert 0 < n_ras < n


def _packed_hand(model, hand):
    import ctypes as Ct
    L = native.lib()
    L.rp_debug_link_table.argtypes = [Ct.c_void_p, Ct.c_int, Ct.POINTER(Ct.c_double)]
    L.rp_debug_geom_table.argtypes = [Ct.c_void_p, Ct.c_int, Ct.POINTER(Ct.c_double)]
    links = np.zeros((26, 28)); geoms = np.zeros((20, 9))
    nl = L.rp_debug_link_table(model.ptr, hand, links.ctypes.data_as(Ct.POINTER(Ct.c_double)))
    ng = L.rp_debug_geom_table(model.ptr, hand, geoms.ctypes.data_as(Ct.POINTER(Ct.c_double)))
    return nl, links, ng, geoms


def _link_kinematics(links, q):
    """World pose of every link slot from the packed tables (what the kernel's level loop computes)."""
    n = links.shape[0]
    pos, quat, anchor, axis = np.zeros((n, 3)), np.zeros((n, 4)), np.zeros((n, 3)), np.zeros((n, 3))
    for l in range(n):
        par = int(links[l, 0])
        pp, pq = (pos[par], quat[par]) if par >= 0 else (np.zeros(3), np.array([1.0, 0, 0, 0]))
        p = pp + mu.rotate(pq, links[l, 3:6])
        qn = mu.quat_mul(pq, links[l, 6:10])
        anchor[l] = p + mu.rotate(qn, links[l, 10:13]); axis[l] = mu.rotate(qn, links[l, 13:16])
        if int(links[l, 2]) == 2:
            p = p + axis[l] * q[l]
        else:
            qn = mu.quat_mul(qn, mu.axis_angle_to_quat(links[l, 13:16], q[l]))
            p = anchor[l] - mu.rotate(qn, links[l, 10:13])
        pos[l], quat[l] = p, mu.quat_normalize(qn)
    return pos, quat, anchor, axis


@pytest.mark.parametrize("name", ["ras", "ras_fd", "hc_ras", "fdx"])
def test_packed_links_reproduce_the_flat_model(name):
    """The C-ABI packer (host side of rp_model_create) against the flat model it was given, at a random in-range pose: bodies
    without joints (reduced action space, a forearm without dofs) are fused into their parents' links — geoms re-expressed,
    inertials merged by an eigen-decomposition — and unused link slots padded with inert links.  The mass matrix assembled from
    the packed links equals the flat model's (bodies kept separate), the packed geoms sit where the flat model's do."""
    kw = VARIANTS[name]
    fm = cache.load_scene(gravity_compensation=True, **kw)
    goal, fing = helpers.song_tables()
    model = native.Model(fm, goal, fing, kind=0, n_steps_lookahead=1)
    rng = np.random.default_rng(8)
    nv = fm.meta["nv"]
    q = fm.jnt_range[:, 0] + rng.random(nv) * (fm.jnt_range[:, 1] - fm.jnt_range[:, 0])
    M = np_ref.mass_matrix(fm, q)
    xpos, xquat, _, _ = np_ref.fk(fm, q)
    for h in range(2):
        nl, links, ng, geoms = _packed_hand(model, h)
        assert nl == fm.arrays["hand_obs_dofid"].shape[1]
        base = 88 + h * nl
        qh = np.zeros(26); qh[:nl] = q[base:base + nl]
        pos, quat, anchor, axis = _link_kinematics(links, qh)
        # padding links: inert one-dof dummies
        for l in range(nl, 26):
            assert links[l, 0] == -1 and links[l, 16] == 0 and links[l, 27] == 1.0
        # mass matrix from the packed links (Jacobian sum over the link inertials)
        chain = [[] for _ in range(26)]
        for l in range(26):
            par = int(links[l, 0])
            chain[l] = (chain[par] if par >= 0 else []) + [l]
        Mp = np.zeros((26, 26))
        for l in range(nl):
            m = links[l, 16]
            if m > 0:
                com = pos[l] + mu.rotate(quat[l], links[l, 17:20])
                jp, jr = np.zeros((3, 26)), np.zeros((3, 26))
                for j in chain[l]:
                    if int(links[j, 2]) == 3:
                        jr[:, j] = axis[j]; jp[:, j] = np.cross(axis[j], com - anchor[j])
                    else:
                        jp[:, j] = axis[j]
                R = mu.quat_to_mat(mu.quat_mul(quat[l], links[l, 20:24]))
                Mp += m * jp.T @ jp + jr.T @ (R @ np.diag(links[l, 24:27]) @ R.T) @ jr
        Mp[np.diag_indices(26)] += links[:, 27]
        assert np.abs(Mp[:nl, :nl] - M[base:base + nl, base:base + nl]).max() < 1e-12
        # packed colliding primitives against the flat model's geoms
        flat = [g for g in range(fm.meta["ngeom"]) if fm.geom_hand[g] == h and fm.geom_collide[g]]
        assert ng == len(flat) == 18
        for gi, g in enumerate(flat):
            l = int(geoms[gi, 0])
            b = int(fm.geom_bodyid[g])
            want = xpos[b] + mu.rotate(xquat[b], fm.geom_pos[g])
            got = pos[l] + mu.rotate(quat[l], geoms[gi, 2:5])
            assert np.abs(got - want).max() < 1e-12, (name, h, gi)
            if int(geoms[gi, 1]) == 3:
                ax_want = mu.rotate(mu.quat_mul(xquat[b], fm.geom_quat[g]), np.array([0.0, 0, 1]))
                assert np.abs(mu.rotate(quat[l], geoms[gi, 5:8]) - ax_want).max() < 1e-12
            assert geoms[gi, 8] == fm.geom_size[g, 0]


@pytest.mark.parametrize("name", ["hc", "hc_ras"])
def test_kernel_pair_list_equals_the_oracles(name):
    """The hand geom pairs the kernel tests (packed by rp_model_create: contype / conaffinity, weld bodies, parent-child,
    <exclude>) are the oracle's, in the same order — MuJoCo's contact order depends on it."""
    import ctypes as Ct
    fm = cache.load_scene(gravity_compensation=True, **VARIANTS[name])
    goal, fing = helpers.song_tables()
    model = native.Model(fm, goal, fing, kind=0, n_steps_lookahead=1)
    L = native.lib()
    L.rp_debug_pair_list.argtypes = [Ct.c_void_p, Ct.POINTER(Ct.c_int32)]
    buf = np.zeros((640, 2), np.int32)
    n = L.rp_debug_pair_list(model.ptr, buf.ctypes.data_as(Ct.POINTER(Ct.c_int32)))
    flat = [[g for g in range(fm.meta["ngeom"]) if fm.geom_hand[g] == h and fm.geom_collide[g]] for h in (0, 1)]
    mine = [(flat[a // 20][a % 20], flat[b // 20][b % 20]) for a, b in buf[:n]]
    om = O.OracleModel(fm)
    OL = O.lib()
    OL.orc_npair.argtypes = [Ct.c_void_p]; OL.orc_pair_get.argtypes = [Ct.c_void_p, Ct.c_int, Ct.POINTER(Ct.c_int), Ct.POINTER(Ct.c_int)]
    theirs = []
    for i in range(OL.orc_npair(om.ptr)):
        a, b = Ct.c_int(), Ct.c_int()
        OL.orc_pair_get(om.ptr, i, Ct.byref(a), Ct.byref(b))
        theirs.append((a.value, b.value))
    assert mine == theirs and n > 400
    # a model without hand collisions has none
    off = native.Model(cache.load_scene(gravity_compensation=True), goal, fing, kind=0, n_steps_lookahead=1)
    assert L.rp_debug_pair_list(off.ptr, buf.ctypes.data_as(Ct.POINTER(Ct.c_int32))) == 0

"""CPU oracle pinned against what can be pinned without MuJoCo: an independent numpy restatement
of kinematics / inertia (finite differences for the bias force), closed-form key statics, the
reference's threshold/structural tests, and committed golden trajectories (regression)."""
import numpy as np
import pytest

import helpers
import np_ref
from oracle import oracle as O
from oracle import task_ref
from robopianist_b200.scene import cache


@pytest.fixture(scope="module")
def hands():
    fm = cache.load_scene(gravity_compensation=True)
    om = O.OracleModel(fm)
    om.set_option(maxcon=24)
    return fm, om


def test_mass_matrix_and_bias_vs_numpy(hands):
    fm, om = hands
    d = O.OracleData(om)
    rng = np.random.default_rng(0)
    lo, hi = fm.jnt_range[:, 0], fm.jnt_range[:, 1]
    q = lo + rng.random(140) * (hi - lo)
    dofs = list(range(88, 100)) + [0]
    v = np.zeros(140); v[dofs] = rng.normal(size=len(dofs))
    d.qpos[:] = q; d.qvel[:] = v
    for fn in ("orc_kinematics", "orc_crb", "orc_velocity"):
        d.call(fn)
    assert np.abs(np_ref.mass_matrix(fm, q) - d.M).max() < 1e-12
    fm0 = cache.load_scene(gravity_compensation=False)
    d0 = O.OracleData(O.OracleModel(fm0))
    d0.qpos[:] = q; d0.qvel[:] = v
    for fn in ("orc_kinematics", "orc_crb", "orc_velocity"):
        d0.call(fn)
    b = np_ref.bias_fd(fm0, q, v, dofs)
    assert np.abs(b[dofs] - d0.qfrc_bias[dofs]).max() < 1e-6


def test_key_rests_on_lower_limit(hands):
    fm, om = hands
    d = O.OracleData(om)
    d.reset()
    for _ in range(400):
        d.mj_step(np.zeros(44))
    # spring (-1 deg reference) beats gravity: keys sit slightly below 0 on the soft limit, at rest
    assert (d.qpos[:88] < 0).all() and (d.qpos[:88] > -2e-3).all() and np.abs(d.qvel[:88]).max() < 1e-6


def test_wrong_press_terminates_in_one_step():
    # piano_with_shadow_hands_test.py:228-242: qfrc_applied = 3 on every key -> LAST, discount 0 after one step
    fm = cache.load_scene()
    goal, fing = helpers.song_tables("CMajorScaleTwoHands", trim_silence=False)
    env = task_ref.RefEnv(fm, goal, fing, wrong_press_termination=True)
    env.reset()
    env.d.qfrc_applied[:88] = 3.0
    st, rew, disc, obs = env.step(np.zeros(45))
    assert st == task_ref.LAST and disc == 0.0


def test_episode_length_and_goal_observable():
    # piano_with_shadow_hands_test.py:119-192: a T-step song ends after exactly T steps; goal obs = roll slice
    fm = cache.load_scene()
    goal, fing = helpers.song_tables("CMajorChordProgressionTwoHands", trim_silence=False)
    goal, fing = goal[:6], fing[:6]
    env = task_ref.RefEnv(fm, goal, fing, n_steps_lookahead=2)
    st, r, dsc, obs = env.reset()
    assert st == task_ref.FIRST and r is None and obs.shape == (88 + 1 + 52 + 3 * 89 + 10,)
    for t in range(6):
        st, r, dsc, obs = env.step(np.zeros(45))
        g = obs[141:141 + 3 * 89].reshape(3, 89)
        if t < 5:
            assert st == task_ref.MID
            for i in range(3):
                exp = goal[t + 1 + i] if t + 1 + i < 6 else np.zeros(89)
                assert (g[i] == exp).all()
    assert st == task_ref.LAST and dsc == 1.0


def test_tolerance_values():
    assert task_ref.tolerance(0.02, (0, 0.05), 0.5) == 1.0
    assert np.isclose(task_ref.tolerance(0.55, (0, 0.05), 0.5), 0.1)   # value_at_margin
    assert np.isclose(task_ref.tolerance(-0.5, (0, 0.05), 0.5), 0.1)


def test_action_replay_presses_the_goal_keys(hands):
    """The reference's recorded policy actions (trained on real MuJoCo) played open loop through the
    oracle still press most goal keys: an external anchor for the restated physics."""
    fm, om = hands
    goal, fing = helpers.song_tables()
    env = task_ref.RefEnv(fm, goal, fing)
    env.reset()
    acts = helpers.twinkle_actions()
    hit = tot = 0
    for t in range(158):
        g = goal[t, :88].astype(bool)
        st, r, dsc, obs = env.step(acts[t])
        hit += (env.activation & g).sum(); tot += g.sum()
    assert st == task_ref.LAST and hit / tot > 0.7


def test_golden_trajectory_regression(hands):
    fm, om = hands
    gold = np.load(helpers.GOLDEN / "oracle_twinkle_traj.npz")
    d = O.OracleData(om)
    d.reset()
    acts = helpers.twinkle_actions()
    for t in range(60):
        d.control_step(acts[t, :44])
        if t in (9, 29, 59):
            assert np.abs(d.qpos - gold[f"qpos_{t}"]).max() < 1e-9


def test_prf_binary_matches_sklearn():
    """The restated per-step metric of MidiEvaluationWrapper (wrappers/evaluation.py:140-147) against
    sklearn itself, including every zero-division corner."""
    from sklearn.metrics import precision_recall_fscore_support
    rng = np.random.default_rng(0)
    cases = [(np.zeros(88), np.zeros(88)), (np.zeros(88), np.ones(88)), (np.ones(88), np.zeros(88)), ([1.0], [0.0]), ([0.0], [0.0])]
    for _ in range(50):
        p = rng.choice([0.02, 0.1, 0.5])
        cases.append(((rng.random(88) < p).astype(float), (rng.random(88) < p).astype(float)))
    for y_true, y_pred in cases:
        ref = precision_recall_fscore_support(y_true=np.asarray(y_true), y_pred=np.asarray(y_pred), average="binary", zero_division=1)[:3]
        assert np.allclose(task_ref.prf_binary(y_true, y_pred), ref, atol=1e-15), (y_true, y_pred)


def test_midi_messages_of_the_reference_restatement(hands):
    """RefEnv(record_midi=True) restates MidiModule.after_substep: every note_on is closed by a note_off of the
    same note unless still held, times are multiples of the physics timestep, pedal events alternate."""
    fm = hands[0]
    goal, fing = helpers.song_tables("TwinkleTwinkleRousseau")
    ref = task_ref.RefEnv(fm, goal, fing, n_steps_lookahead=1)
    ref.record_midi = True
    ref.reset()
    acts = helpers.twinkle_actions()
    acts[:, -1] = (np.arange(len(acts)) // 7) % 2
    for t in range(45):
        ref.step(acts[t])
    msgs = ref.midi_messages
    assert len(msgs) > 10
    held, pedal = set(), False
    for tm, kind, note in msgs:
        assert abs(tm / 0.005 - round(tm / 0.005)) < 1e-9
        if kind == "note_on":
            assert note not in held; held.add(note)
        elif kind == "note_off":
            assert note in held; held.remove(note)
        elif kind == "sustain_on":
            assert not pedal; pedal = True
        else:
            assert pedal; pedal = False
    assert [m[0] for m in msgs] == sorted(m[0] for m in msgs)
    # identical physics with and without per-substep stepping
    ref2 = task_ref.RefEnv(fm, goal, fing, n_steps_lookahead=1)
    ref2.reset()
    for t in range(45):
        ref2.step(acts[t])
    assert np.array_equal(ref.d.qpos, ref2.d.qpos)


def test_pgs_fixed_point_satisfies_the_kkt_conditions(hands):
    """An implementation-independent property of the solve: run to convergence, the PGS forces must satisfy the
    optimality conditions of MuJoCo's dual problem min 1/2 f'(A+R)f + f'b over the constraint set — friction-loss
    rows inside their box with the gradient pointing out of it at an active bound, limit rows f >= 0, g >= 0,
    f g = 0, elliptic contacts f in the friction cone, g in its dual cone, f . g = 0."""
    fm, _ = hands
    om = O.OracleModel(fm, iterations=4000, tolerance=0.0)
    om.set_option(maxcon=24)
    d = O.OracleData(om)
    d.reset()
    acts = helpers.twinkle_actions()[:, :44]
    checked = 0
    for t in range(26):
        d.control_step(acts[t])
        if t < 16 or d.ncon == 0:
            continue
        d.ctrl[:] = acts[t + 1]
        d.call("orc_smooth"); d.call("orc_solve_init"); d.call("orc_solve_iterate")
        typ, eid = d.efc_type(), d.efc_id()
        AR, b, f, loss = d.efc("efc_AR"), d.efc("efc_b"), d.efc("efc_force"), d.efc("efc_frictionloss")
        g = AR @ f + b
        sc = 1e-7 * max(1.0, np.abs(b).max())
        cons = d.contacts()
        i = 0
        while i < len(typ):
            if typ[i] == 0:      # friction loss: box [-loss, loss]
                assert abs(f[i]) <= loss[i] + 1e-12
                if abs(f[i]) < loss[i] - 1e-9:
                    assert abs(g[i]) < sc
                else:
                    assert g[i] * np.sign(f[i]) < sc
                i += 1
            elif typ[i] == 1:    # limit: f >= 0 complementary to g >= 0
                assert f[i] >= 0 and g[i] > -sc and abs(f[i] * g[i]) < sc * max(1.0, f[i])
                i += 1
            else:                # elliptic contact, condim 3
                mu = cons[eid[i]]["friction"]   # the dual solver works on the contact's own friction cone
                fn, ft = f[i], np.hypot(f[i + 1], f[i + 2])
                gn, gt = g[i], np.hypot(g[i + 1], g[i + 2])
                assert fn >= 0 and ft <= mu * fn + 1e-9
                assert gn > -sc and mu * gt <= gn + 10 * sc
                assert abs(f[i:i + 3] @ g[i:i + 3]) < 10 * sc * max(1.0, fn)
                i += 3
        checked += 1
        # leave the state as control_step expects it (the next step2 recomputes the solve)
    assert checked >= 5


@pytest.mark.parametrize("key", [40, 42])
def test_key_equilibrium_under_a_constant_torque_is_the_analytic_one(hands, key):
    """Closed-form statics of one key (piano_mjcf.py:66-100: hinge about +y at the back edge, spring towards
    -1 deg, uniform box): with a constant generalized force tau the rest angle solves
    tau + m g (r_x cos q + r_z sin q) - k (q - q_ref) = 0 (gravity torque about the hinge of the rotated
    centre of mass); the oracle, stepped to rest, must land on it (no limit active in between)."""
    fm, om = hands
    A = fm.arrays
    dof, body = int(A["key_dofid"][key]), int(A["key_bodyid"][key])
    jnt = dof  # one hinge per key, joints and dofs share the numbering
    m, k, qref = A["body_mass"][body], A["jnt_stiffness"][jnt], A["jnt_springref"][jnt]
    r0 = A["body_ipos"][body] - A["jnt_pos"][jnt]
    assert np.allclose(A["jnt_axis"][jnt], [0, 1, 0])
    g = -A["opt_gravity"][2]
    qmax = fm.key_qmax[key]
    target = 0.45 * qmax
    grav = lambda q: m * g * (r0[0] * np.cos(q) + r0[2] * np.sin(q))   # generalized gravity force about +y
    tau = k * (target - qref) - grav(target)
    d = O.OracleData(om)
    d.reset()
    d.qfrc_applied[dof] = tau
    for _ in range(800):
        d.mj_step(np.zeros(44))
    assert abs(d.qvel[dof]) < 1e-9
    # the bias force the oracle computes at that angle is minus the analytic gravity torque
    assert abs(d.qfrc_bias[dof] + grav(d.qpos[dof])) < 1e-12
    assert abs(d.qpos[dof] - target) < 1e-8
    # the neighbours are untouched and still resting on their lower limit
    assert d.qpos[dof - 1] < 0 and d.qpos[dof + 1] < 0


def _point_box_distance(p, half):
    q = np.abs(p) - half
    return np.linalg.norm(np.maximum(q, 0.0)) + min(q.max(), 0.0)   # signed: negative inside


def test_capsule_box_contacts_are_geometrically_consistent(hands):
    """Narrowphase output against brute force, independent of how it was computed: for every capsule / sphere -
    box pair in contact, (i) the deepest reported distance equals the minimum over the capsule axis of the
    signed point-box distance minus the radius (dense sampling), (ii) every contact's normal is a unit vector
    and its position is the mid-point between a point on the box surface and the surface point of a sphere of
    the capsule's radius centred on the capsule axis, `dist` apart along the normal."""
    fm, om = hands
    A = fm.arrays
    acts = helpers.twinkle_actions()[:, :44]
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    rng = np.random.default_rng(4)
    pairs = contacts = 0
    for e in range(3):
        d = O.OracleData(om)
        d.reset()
        for t in range(30):
            d.control_step(np.clip(acts[t] + (rng.normal(0, 0.08, 44) * (hi - lo) / 2 if e else 0), lo, hi))
            if t < 12:
                continue
            by_pair = {}
            for c in d.contacts():
                by_pair.setdefault((c["geom1"], c["geom2"]), []).append(c)
            for (g1, g2), cs in by_pair.items():
                assert A["geom_type"][g2] == 6 and A["geom_type"][g1] in (2, 3)   # sphere / capsule vs box
                R2 = d.geom_xmat[g2].reshape(3, 3); p2 = d.geom_xpos[g2]; half = A["geom_size"][g2]
                R1 = d.geom_xmat[g1].reshape(3, 3); p1 = d.geom_xpos[g1]
                rad = A["geom_size"][g1][0]
                hl = A["geom_size"][g1][1] if A["geom_type"][g1] == 3 else 0.0
                ts = np.linspace(-hl, hl, 4001 if hl > 0 else 1)
                axis_pts = p1[None, :] + ts[:, None] * R1[:, 2][None, :]
                local = (axis_pts - p2) @ R2                                  # box frame
                dmin = min(_point_box_distance(q, half) for q in local) - rad
                assert abs(min(c["dist"] for c in cs) - dmin) < 2e-7, (t, g1, g2)
                for c in cs:
                    n = c["frame"][0]
                    assert abs(np.linalg.norm(n) - 1) < 1e-12
                    s1 = c["pos"] - 0.5 * c["dist"] * n      # on the capsule surface
                    s2 = c["pos"] + 0.5 * c["dist"] * n      # on the box surface
                    assert abs(_point_box_distance((s2 - p2) @ R2, half)) < 1e-9
                    centre = s1 - rad * n                    # the sphere this contact was generated from
                    rel = centre - p1
                    tt = np.clip(rel @ R1[:, 2], -hl, hl)
                    assert np.linalg.norm(rel - tt * R1[:, 2]) < 1e-9   # ... sits on the capsule axis segment
                    contacts += 1
                pairs += 1
    assert pairs > 40 and contacts > 60


def test_contact_jacobian_matches_finite_differences(hands):
    """Rows of efc_J for contacts against finite differences of the kinematics: moving dof j by eps displaces the
    material points of the two bodies at the contact position; (displacement of body 2 - body 1) / eps along
    the three frame axes is column j of the contact's Jacobian block."""
    fm, om = hands
    A = fm.arrays
    acts = helpers.twinkle_actions()[:, :44]
    d = O.OracleData(om); d.reset()
    probe = O.OracleData(om)
    eps, checked = 1e-7, 0
    for t in range(26):
        d.control_step(acts[t])
        if t < 17 or d.ncon == 0:
            continue
        typ, eid, J = d.efc_type(), d.efc_id(), d.efc("efc_J")
        cons = d.contacts()
        rows = [i for i in range(len(typ)) if typ[i] == 2][::3]
        xpos0, xmat0 = d.xpos.copy(), d.xmat.copy()
        for i in rows[:3]:
            c = cons[eid[i]]
            b1, b2 = int(A["geom_bodyid"][c["geom1"]]), int(A["geom_bodyid"][c["geom2"]])
            loc = [xmat0[b].reshape(3, 3).T @ (c["pos"] - xpos0[b]) for b in (b1, b2)]
            cols = np.flatnonzero(np.abs(J[i:i + 3]).sum(axis=0))
            assert len(cols) >= 5
            for j in cols:
                probe.qpos[:] = d.qpos
                probe.qpos[j] += eps
                probe.call("orc_kinematics")
                p = [probe.xpos[b] + probe.xmat[b].reshape(3, 3) @ l for b, l in zip((b1, b2), loc)]
                num = c["frame"] @ ((p[1] - p[0]) / eps)      # both were at c["pos"] before the perturbation
                assert np.abs(num - J[i:i + 3, j]).max() < 5e-6, (t, i, j, num, J[i:i + 3, j])
            # a dof outside the chain moves neither body
            other = [j for j in range(fm.meta["nv"]) if j not in cols][:3]
            assert np.abs(J[i:i + 3][:, other]).max() == 0
            checked += 1
    assert checked >= 8


def test_step_invariants_recomputed_in_numpy(hands):
    """One mj_step of the oracle re-derived from its own intermediate arrays with dense numpy algebra:
    actuator forces from the position-actuator model (clamped ctrl, kp (ctrl - length), force range), the
    dual -> primal map M (qacc - qacc_smooth) = J^T f, the implicit-damping Euler system
    (M + h D)(v+ - v) / h = qfrc_smooth + qfrc_constraint, and q+ = q + h v+."""
    fm, om = hands
    A = fm.arrays
    h = fm.meta["timestep"]
    acts = helpers.twinkle_actions()[:, :44]
    d = O.OracleData(om); d.reset()
    checked = 0
    for t in range(24):
        d.control_step(acts[t])
        if t < 15:
            continue
        q0, v0 = d.qpos.copy(), d.qvel.copy()
        ctrl = acts[t + 1]
        d.ctrl[:] = ctrl
        d.step2()      # d holds a valid step1 for (q0, v0); step2 = actuation .. Euler
        M = d.M.copy()
        # actuators: joint / fixed-tendon transmission, gain kp, bias -kp * length
        length = d.actuator_length.copy()
        kp = A["actuator_gainprm"]
        c = np.where(A["actuator_ctrllimited"].astype(bool), np.clip(ctrl, A["actuator_ctrlrange"][:, 0], A["actuator_ctrlrange"][:, 1]), ctrl)
        force = kp * c + A["actuator_biasprm"][:, 0] + A["actuator_biasprm"][:, 1] * length + A["actuator_biasprm"][:, 2] * d.actuator_velocity
        force = np.where(A["actuator_forcelimited"].astype(bool), np.clip(force, A["actuator_forcerange"][:, 0], A["actuator_forcerange"][:, 1]), force)
        # (step2 ends with Euler: actuator_length / velocity still belong to (q0, v0), the state the force was computed from)
        assert np.abs(force - d.actuator_force).max() < 1e-12
        J, f = d.efc("efc_J"), d.efc("efc_force")
        assert np.abs(M @ (d.qacc - d.qacc_smooth) - J.T @ f).max() < 1e-9
        assert np.abs(J.T @ f - d.qfrc_constraint).max() < 1e-12
        D = np.diag(A["dof_damping"])
        lhs = (M + h * D) @ ((d.qvel - v0) / h)
        assert np.abs(lhs - (d.qfrc_smooth + d.qfrc_constraint)).max() < 1e-8
        assert np.abs(d.qpos - (q0 + h * d.qvel)).max() < 1e-15
        d.step1()      # restore the invariant control_step relies on
        checked += 1
    assert checked >= 8


def test_ot_fingering_reward_of_the_reference_restatement(hands):
    """piano_with_shadow_hands.py:130-141: with the fingering reward disabled (or no fingering in the song) the task
    adds the optimal-transport fingering reward; RefEnv(ot_fingering_reward=True) restates it with scipy's
    linear_sum_assignment, as the reference does: term in (0, 1], exactly 1 when no key is to be pressed, part of
    the reward, and not larger than the assignment-free upper bound (every key matched to its nearest fingertip)."""
    fm, _ = hands
    goal, fing = helpers.song_tables("TwinkleTwinkleRousseau")
    ref = task_ref.RefEnv(fm, goal, -np.ones_like(fing), disable_fingering_reward=True, ot_fingering_reward=True)
    ref.reset()
    acts = helpers.twinkle_actions()
    saw_one = saw_less = 0
    for t in range(40):
        gprev = ref.goal_state[0].copy()
        st, rew, disc, obs = ref.step(acts[t])
        terms = ref.reward_terms
        assert set(terms) == {"key_press_reward", "sustain_reward", "energy_reward", "ot_fingering_reward", "forearm_reward"}
        assert abs(rew - sum(terms.values())) < 1e-12
        r = terms["ot_fingering_reward"]
        keys = np.flatnonzero(gprev[:-1])
        if keys.size == 0:
            assert r == 1.0; saw_one += 1
        else:
            assert 0.0 < r <= 1.0; saw_less += r < 1.0
            d = ref.d
            kp = np.array([d.geom_xpos[fm.key_geomid[k]] + [0.35 * fm.geom_size[fm.key_geomid[k], 0], 0, 0.5 * fm.geom_size[fm.key_geomid[k], 2]] for k in keys])
            nearest = np.linalg.norm(kp[:, None, :] - d.site_xpos[None, :10, :], axis=2).min(axis=1)
            assert r <= float(np.mean(task_ref.tolerance(nearest, (0, 0.01), 0.1))) + 1e-12
    assert saw_one > 0 and saw_less > 0
    assert obs.shape == (88 + 1 + 52 + 2 * 89,)   # no fingering observable


def test_warm_start_uses_the_regularised_cone_and_the_fixed_point_does_not_care(hands):
    """mj_makeImpedance gives the friction rows of an elliptic contact R_t = R_n / impratio and the primal map of
    mj_constraintUpdate the friction coefficient of the *regularised* cone, friction * sqrt(R_t / R_n); the dual
    solver keeps the contact's own friction.  (1) The oracle's warm-start forces are the numpy restatement of that
    map and differ from the un-regularised variant (impratio = 10 here) on sliding contacts; (2) every warm-start
    force lies in the friction cone; (3) the converged PGS forces do not depend on the starting point."""
    import ctypes as C
    fm, _ = hands
    om = O.OracleModel(fm, iterations=6000, tolerance=0.0)
    om.set_option(maxcon=24)
    d = O.OracleData(om)
    d.reset()
    acts = helpers.twinkle_actions()[:, :44]
    nv = om.nv

    def primal(jar, D, mu, fr):
        N, T = jar[0] * mu, fr * np.hypot(jar[1], jar[2])
        if N >= mu * T or (T <= 0 and N >= 0):
            return np.zeros(3)
        if mu * N + T <= 0 or (T <= 0 and N < 0):
            return -D * jar
        f0 = -D[0] / (mu * mu * (1 + mu * mu)) * (N - mu * T) * mu
        return np.array([f0, -f0 / T * jar[1] * fr * fr, -f0 / T * jar[2] * fr * fr])

    differs = middle = checked = 0
    for t in range(40):
        d.control_step(acts[t])
        if d.ncon == 0:
            continue
        d.ctrl[:] = acts[t + 1]
        d.call("orc_smooth"); d.call("orc_solve_init")
        typ, eid = d.efc_type(), d.efc_id()
        J, aref, R, f_ws = d.efc("efc_J"), d.efc("efc_aref"), d.efc("efc_R"), d.efc("efc_force")
        jar = J @ d.qacc_warmstart - aref
        cons = d.contacts()
        kept = np.abs(f_ws).max() > 0          # the warm start is dropped as a whole when its dual cost is positive
        i = 0
        while i < len(typ):
            if typ[i] != 2:
                i += 1
                continue
            fr = cons[eid[i]]["friction"]
            assert abs(R[i + 1] * fm.meta["impratio"] - R[i]) < 1e-12 * R[i]
            mu = fr * np.sqrt(R[i + 1] / R[i])
            assert abs(cons[eid[i]]["mu"] - mu) < 1e-15
            want = primal(jar[i:i + 3], 1 / R[i:i + 3], mu, fr)
            other = primal(jar[i:i + 3], 1 / R[i:i + 3], fr, fr)
            if kept:
                assert np.allclose(f_ws[i:i + 3], want, rtol=1e-10, atol=1e-12)
                assert np.hypot(want[1], want[2]) <= fr * want[0] * (1 + 1e-9) + 1e-12
                if want[0] > 0 and abs(np.hypot(want[1], want[2]) - fr * want[0]) < 1e-9 * want[0]:
                    middle += 1
                differs += int(not np.allclose(want, other, rtol=1e-6, atol=1e-9))
            i += 3
        if checked < 4:
            # converged forces from the warm start and from zero
            d.call("orc_solve_iterate")
            f_a = d.efc("efc_force")
            d.call("orc_smooth"); d.call("orc_solve_init")
            p = O.lib().orc_ptr(d.ptr, b"efc_force")
            np.ctypeslib.as_array(p, shape=(d.nefc,))[:] = 0
            d.call("orc_solve_iterate")
            f_b = d.efc("efc_force")
            assert np.abs(f_a - f_b).max() < 1e-6 * max(1.0, np.abs(f_a).max())
            checked += 1
    assert checked == 4 and differs > 0 and middle > 0


def _surface_points(gtype, size):
    if gtype in (6, 7):      # box, or the bounding box kept for a mesh
        u = np.linspace(-1, 1, 7)
        pts = []
        for ax in range(3):
            a, b = [k for k in range(3) if k != ax]
            U, V = np.meshgrid(u, u, indexing="ij")
            for sgn in (-1, 1):
                p = np.zeros((U.size, 3)); p[:, a] = U.ravel() * size[a]; p[:, b] = V.ravel() * size[b]; p[:, ax] = sgn * size[ax]
                pts.append(p)
        return np.concatenate(pts)
    r, hh = size[0], size[1]   # cylinder, axis z
    th = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    side = np.array([[r * np.cos(t), r * np.sin(t), z] for z in np.linspace(-hh, hh, 5) for t in th])
    caps = np.array([[rr * np.cos(t), rr * np.sin(t), s * hh] for s in (-1, 1) for rr in (0, r / 2, r) for t in th])
    return np.concatenate([side, caps])


def test_uncovered_hand_colliders_stay_clear_of_the_piano_on_the_replay(hands):
    """40 hand colliders of the reference model (12 boxes, 7 cylinders and the forearm mesh per hand:
    right_hand.xml:101-288) are outside the narrowphase's coverage (meta['ignored_colliders']; MuJoCo's box-box and
    general convex routines are not restated).  On the recorded action replay they never come near the piano: brute
    force, surface samples of every such geom against every key / base box, every fourth control step, clearance
    > 3 cm.  (At runtime the kernel flags any environment in which one of them comes within reach: diag bit 3.)"""
    fm, _ = hands
    names = fm.meta["geom_names"]
    ign = [names.index(n) for n in fm.meta["ignored_colliders"]]
    assert len(ign) == fm.meta["n_ignored_colliders"] == 40
    assert sorted(int(fm.geom_type[g]) for g in ign) == [5] * 14 + [6] * 24 + [7] * 2
    assert not fm.geom_collide[ign].any()
    piano = [i for i in range(len(names)) if fm.geom_key[i] != -1]
    P = {g: _surface_points(int(fm.geom_type[g]), fm.geom_size[g]) for g in ign}
    om = O.OracleModel(fm); om.set_option(maxcon=24)
    d = O.OracleData(om); d.reset()
    acts = helpers.twinkle_actions()
    half = fm.geom_size[piano]
    best = 1e9
    for t in range(acts.shape[0]):
        d.control_step(acts[t, :44])
        if t % 4:
            continue
        c = d.geom_xpos[piano]; R = d.geom_xmat[piano].reshape(-1, 3, 3)
        for g in ign:
            pw = d.geom_xpos[g] + P[g] @ d.geom_xmat[g].reshape(3, 3).T
            loc = np.einsum("pkc,kcd->pkd", pw[:, None, :] - c[None], R)
            q = np.abs(loc) - half[None]
            sd = np.linalg.norm(np.maximum(q, 0), axis=2) + np.minimum(q.max(axis=2), 0)
            best = min(best, sd.min())
    assert best > 0.03, best


def _uncovered_boxes_near_piano(fm, d, q):
    """The kernel's runtime test (rp_kernels.cu, diag flag 8) restated: world AABB of each uncovered collider's oriented
    bounding box against the piano's bounding box."""
    names = fm.meta["geom_names"]
    ign = [names.index(n) for n in fm.meta["ignored_colliders"]]
    piano = [i for i in range(len(names)) if fm.geom_key[i] != -1]
    d.qpos[:] = 0; d.call("orc_kinematics")
    lo = np.min(d.geom_xpos[piano] - fm.geom_size[piano], axis=0); hi = np.max(d.geom_xpos[piano] + fm.geom_size[piano], axis=0)
    d.qpos[:] = q; d.call("orc_kinematics")
    out = []
    for g in ign:
        h = fm.geom_size[g].copy()
        if fm.geom_type[g] == 5:
            h = np.array([h[0], h[0], h[1]])
        e = np.abs(d.geom_xmat[g].reshape(3, 3)) @ h
        c = d.geom_xpos[g]
        if np.all(c + e > lo) and np.all(c - e < hi):
            out.append(names[g])
    return out


def test_uncovered_colliders_cannot_reach_the_piano_within_joint_limits(hands):
    """Why leaving the palm / wrist / knuckle / forearm colliders out of the piano narrowphase changes nothing while the
    joints respect their ranges: the hands hang at x = 0.4 m pointing at the keys (tasks/base.py:34-37, 173-183), the
    forearm slides only sideways and 6 cm in depth (shadow_hand.py:42-52), and every one of those colliders stays
    outside the piano's bounding box for any joint configuration inside the limits (4000 random poses); a pose far
    outside them (forearm 15 cm in, wrist bent 90 degrees) does reach it — that is what the runtime flag reports."""
    fm, _ = hands
    om = O.OracleModel(fm); d = O.OracleData(om); d.reset()
    rng = np.random.default_rng(0)
    r = fm.jnt_range
    for _ in range(4000):
        q = np.zeros(140)
        q[88:] = rng.uniform(r[88:, 0], r[88:, 1])
        q[:88] = rng.uniform(r[:88, 0], r[:88, 1])
        assert not _uncovered_boxes_near_piano(fm, d, q)
    q = np.zeros(140); q[89] = 0.15; q[91] = 1.57
    assert "rh_shadow_hand/rh_palm_geom1" in _uncovered_boxes_near_piano(fm, d, q)


# ------------------------------------------------------------------------------------------------------------------
# hand geoms colliding with each other (disable_hand_collisions=False)
@pytest.fixture(scope="module")
def hands_hc():
    fm = cache.load_scene(gravity_compensation=True, disable_hand_collisions=False)
    om = O.OracleModel(fm)
    om.set_option(maxcon=60)
    return fm, om


def _hc_rollout(fm, om, substeps, seed=5, amp=0.6):
    """Hands driven into each other with slowly varying finger targets: cross-hand, same-hand and piano contacts."""
    d = O.OracleData(om); d.reset()
    names = fm.meta["actuator_names"]
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    rng = np.random.default_rng(seed); target = rng.random(fm.meta["nu"])
    for t in range(substeps):
        if t % 100 == 0:
            target = rng.random(fm.meta["nu"])
        if t % 10 == 0:
            ctrl = lo + (hi - lo) * (0.5 + amp * (target - 0.5))
            for i, n in enumerate(names):
                if n.endswith("forearm_tx"):
                    ctrl[i] = (-0.2 if "rh_" in n else 0.2) * min(1.0, t / 200.0)
                if n.endswith("forearm_ty"):
                    ctrl[i] = 0.0
        d.mj_step(ctrl)
        yield t, d


def test_hand_pair_contacts_geometry_and_jacobian(hands_hc):
    """Hand : hand contacts of the oracle, independent of how they were computed: the reported distance is the brute-force
    distance between the two capsule segments minus the radii, the normal points from geom 1 to geom 2 (sphere before capsule),
    the position lies mid-way between the surfaces; the three Jacobian rows equal finite differences of the kinematics
    (body 2 minus body 1), and for two bodies of one hand the columns of their common ancestors vanish exactly — the property
    the kernel's single-record representation of such contacts rests on."""
    fm, om = hands_hc
    A = fm.arrays
    probe = O.OracleData(om)
    eps = 1e-7
    seen = {"cross": 0, "self": 0}
    ts = np.linspace(-1, 1, 2001)
    for t, d in _hc_rollout(fm, om, 460):
        if t % 45 != 44:
            continue
        d.call("orc_step1")   # mj_step leaves the arrays of the state before integration: rebuild them at the current qpos
        if d.ncon == 0:
            continue
        typ, eid, J = d.efc_type(), d.efc_id(), d.efc("efc_J")
        cons = d.contacts()
        rows = [i for i in range(len(typ)) if typ[i] == 2][::3]
        xpos0, xmat0 = d.xpos.copy(), d.xmat.copy()
        for i in rows:
            c = cons[eid[i]]
            g1, g2 = c["geom1"], c["geom2"]
            if A["geom_hand"][g1] < 0 or A["geom_hand"][g2] < 0:
                continue
            kind = "cross" if A["geom_hand"][g1] != A["geom_hand"][g2] else "self"
            if seen[kind] >= 6:
                continue
            assert A["geom_type"][g1] <= A["geom_type"][g2]          # mj_collideGeoms orders the pair by type
            # --- geometry
            seg = []
            for g in (g1, g2):
                hl = A["geom_size"][g][1] if A["geom_type"][g] == 3 else 0.0
                seg.append(d.geom_xpos[g][None, :] + ts[:, None] * hl * d.geom_xmat[g].reshape(3, 3)[:, 2][None, :])
            d2 = ((seg[0][:, None, :] - seg[1][None, :, :]) ** 2).sum(-1)
            ia, ib = np.unravel_index(np.argmin(d2), d2.shape)
            dmin = np.sqrt(d2[ia, ib]) - A["geom_size"][g1][0] - A["geom_size"][g2][0]
            n = c["frame"][0]
            assert abs(c["dist"] - dmin) < 2e-6 and abs(np.linalg.norm(n) - 1) < 1e-12
            assert n @ (seg[1][ib] - seg[0][ia]) > 0                   # from geom 1 towards geom 2
            assert np.abs(c["pos"] - (seg[0][ia] + n * (A["geom_size"][g1][0] + 0.5 * c["dist"]))).max() < 1e-4
            # --- Jacobian: finite differences of the two bodies' material points at the contact position
            b1, b2 = int(A["geom_bodyid"][g1]), int(A["geom_bodyid"][g2])
            loc = [xmat0[b].reshape(3, 3).T @ (c["pos"] - xpos0[b]) for b in (b1, b2)]
            chain = lambda b: set(om_chain(fm, b))
            c1, c2 = chain(b1), chain(b2)
            for j in sorted(c1 | c2):
                probe.qpos[:] = d.qpos
                probe.qpos[j] += eps
                probe.call("orc_kinematics")
                p = [probe.xpos[b] + probe.xmat[b].reshape(3, 3) @ l for b, l in zip((b1, b2), loc)]
                num = c["frame"] @ ((p[1] - p[0]) / eps)
                assert np.abs(num - J[i:i + 3, j]).max() < 5e-6, (t, i, j)
            common = c1 & c2
            if kind == "self":
                assert len(common) >= 4                                # forearm slides + wrist hinges at least
                assert np.abs(J[i:i + 3][:, sorted(common)]).max() == 0.0
                assert len((c1 | c2) - common) <= 10                   # what one contact record of the kernel holds
            else:
                assert not common
            off = [j for j in range(fm.meta["nv"]) if j not in (c1 | c2)][:4]
            assert np.abs(J[i:i + 3][:, off]).max() == 0
            seen[kind] += 1
    assert seen["cross"] >= 3 and seen["self"] >= 3, seen


def om_chain(fm, b):
    """dofs of body b's kinematic chain (root to b)."""
    out = []
    while b > 0:
        out = [int(fm.body_dofadr[b]) + k for k in range(int(fm.body_dofnum[b]))] + out
        b = int(fm.body_parentid[b])
    return out


def test_pgs_fixed_point_with_hand_contacts_satisfies_the_kkt_conditions(hands_hc):
    """The optimality conditions of the dual problem (as in test_pgs_fixed_point_satisfies_the_kkt_conditions) with cross-hand
    and same-hand contacts among the rows: elliptic contacts f in the friction cone, g in its dual cone, f . g = 0."""
    fm, _ = hands_hc
    om = O.OracleModel(fm, iterations=6000, tolerance=0.0)
    om.set_option(maxcon=60)
    checked = pairs = 0
    for t, d in _hc_rollout(fm, om, 360):
        if t % 60 != 59:
            continue
        d.call("orc_step1")
        if d.ncon == 0:
            continue
        d.call("orc_smooth"); d.call("orc_solve_init"); d.call("orc_solve_iterate")
        typ, eid = d.efc_type(), d.efc_id()
        AR, b, f, loss = d.efc("efc_AR"), d.efc("efc_b"), d.efc("efc_force"), d.efc("efc_frictionloss")
        g = AR @ f + b
        sc = 1e-6 * max(1.0, np.abs(b).max())
        cons = d.contacts()
        i = 0
        while i < len(typ):
            if typ[i] == 0:
                assert abs(f[i]) <= loss[i] + 1e-12
                i += 1
            elif typ[i] == 1:
                assert f[i] >= 0 and g[i] > -sc
                i += 1
            else:
                c = cons[eid[i]]
                mu = c["friction"]
                fn, ft = f[i], np.hypot(f[i + 1], f[i + 2])
                gn, gt = g[i], np.hypot(g[i + 1], g[i + 2])
                assert fn >= 0 and ft <= mu * fn + 1e-9
                assert gn > -sc and mu * gt <= gn + 10 * sc
                assert abs(f[i:i + 3] @ g[i:i + 3]) < 10 * sc * max(1.0, fn)
                pairs += int(fm.geom_hand[c["geom1"]] >= 0 and fm.geom_hand[c["geom2"]] >= 0)
                i += 3
        checked += 1
    assert checked >= 4 and pairs >= 4

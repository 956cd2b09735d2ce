"""GPU parity with hand geoms colliding with each other (disable_hand_collisions=False, the reference's default):
sphere / capsule pairs within a hand and across the hands (MuJoCo's analytic routines, restated in the oracle and the
kernel), contact rows over two kinematic chains, the cross-hand Gauss-Seidel coupling, the forearm reward."""
import numpy as np
import pytest

import helpers
from oracle import oracle as O
from oracle import task_ref
from robopianist_b200.scene import cache

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

HC = dict(disable_hand_collisions=False)


def _dev(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def _controls(fm, t, rng_ctrl):
    """Hands driven towards each other (forearm_tx) while the fingers follow slowly varying random targets."""
    names = fm.meta["actuator_names"]
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    ctrl = lo + (hi - lo) * (0.5 + 0.6 * (rng_ctrl - 0.5))  # (full-range targets need more than the 21 contact records)
    for i, n in enumerate(names):
        if n.endswith("forearm_tx"):
            ctrl[i] = (-0.2 if "rh_" in n else 0.2) * min(1.0, t / 200.0)
        if n.endswith("forearm_ty"):
            ctrl[i] = 0.0
    return ctrl


@pytest.mark.parametrize("block_envs,extra", [(1, {}), (3, {}), (3, dict(reduced_action_space=True))])
def test_hand_hand_and_self_contacts_match_oracle(block_envs, extra):
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=3, scene_kw=dict(HC, **extra))
    batch.set_block_envs(block_envs)
    om = O.OracleModel(fm); om.set_option(maxcon=40)
    od = O.OracleData(om); od.reset()
    nv, nu = fm.meta["nv"], fm.meta["nu"]
    batch.set_state(np.zeros((3, nv)), np.zeros((3, nv)), np.zeros((3, nv)))
    od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
    rng = np.random.default_rng(5)
    target = rng.random(nu)
    worst_q = worst_rel = 0.0
    kinds = {"cross": 0, "self": 0, "piano": 0}
    checked = 0
    for t in range(700):
        if t % 100 == 0:
            target = rng.random(nu)
        if t % 10 == 0:
            ctrl = _controls(fm, t, target)
            ct = _dev(np.tile(ctrl, (3, 1)))
        if t % 50 == 0 and t:
            # re-synchronise: with a dozen sliding finger : finger contacts and the sweep cap reached, round-off differences
            # (1e-16 per operation) grow by e^(0.03 t) — 1e-7 after 700 substeps — which would hide nothing but also
            # prove nothing; every 50-substep segment starts from the oracle's state and has to stay within 1e-9
            batch.set_state(np.tile(od.qpos, (3, 1)), np.tile(od.qvel, (3, 1)), np.tile(od.qacc_warmstart, (3, 1)))
        batch.physics_step(ct.data_ptr(), 1)
        od.mj_step(ctrl)
        q, v, _, _ = batch.get_state()
        worst_q = max(worst_q, np.abs(q[0] - od.qpos).max())
        worst_rel = max(worst_rel, (np.abs(v[0] - od.qvel) / (1e-3 + np.abs(od.qvel))).max())
        for c in od.contacts():
            h1, h2 = fm.geom_hand[c["geom1"]], fm.geom_hand[c["geom2"]]
            kinds["piano" if h2 < 0 or h1 < 0 else ("cross" if h1 != h2 else "self")] += 1
        if t % 20 == 19:
            dg = batch.get_diag()
            assert dg["ncon"][0] == od.ncon and dg["niter"][0] == od.solver_niter and (dg["flags"][0] & 7) == 0, (t, dg["ncon"][0], od.ncon, dg["niter"][0], od.solver_niter, dg["flags"][0])
            checked += 1
        assert (q[0] == q[2]).all()
    assert kinds["cross"] > 50 and kinds["self"] > 50 and kinds["piano"] > 50, kinds   # the rollout exercises every kind of pair
    assert worst_q < 1e-9 and worst_rel < 1e-5, (worst_q, worst_rel)
    batch.close()


def test_forearm_reward_and_task_layer_with_hand_collisions():
    """Full step against the restated task on the hand-collision model; an action that drives the forearms into each other
    must zero the forearm reward (piano_with_shadow_hands.py:251-259), the recorded replay must leave it at 0.5."""
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2, scene_kw=HC)
    goal, fing = helpers.song_tables()
    ref = task_ref.RefEnv(fm, goal, fing, n_steps_lookahead=1, maxcon=40)
    N = 2
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda"); terms = torch.empty((N, 5), dtype=torch.float64, device="cuda")
    batch.reset(None, obs.data_ptr())
    ref.reset()
    acts = helpers.twinkle_actions()
    names = fm.meta["actuator_names"]
    seen = set()
    for t in range(70):
        a = acts[t].copy()
        if t >= 25:  # cross the forearms
            for i, n in enumerate(names):
                if n.endswith("forearm_tx"):
                    a[i] = -0.35 if "rh_" in n else 0.35
        batch.step(_dev(np.tile(a, (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr(), terms.data_ptr())
        rst, rrew, rdis, robs = ref.step(a)
        torch.cuda.synchronize()
        assert int(st[0]) == rst
        assert np.abs(obs[0].cpu().numpy() - robs).max() < 1e-9, t
        tr = list(ref.reward_terms.values())
        assert np.abs(terms[0].cpu().numpy()[:len(tr)] - np.array(tr)).max() < 1e-9, (t, terms[0].cpu().numpy(), tr)
        assert abs(rew[0].item() - rrew) < 1e-9
        seen.add(ref.reward_terms["forearm_reward"])
    assert seen == {0.0, 0.5}, seen
    batch.close()


def test_suite_load_with_the_reference_default_hand_collisions():
    """suite.load(name, task_kwargs=dict(primitive_fingertip_collisions=True)): hand collisions on, as in the reference."""
    from robopianist_b200 import suite
    env = suite.load("RoboPianist-debug-TwinkleTwinkleRousseau-v0", seed=0, num_envs=4,
                     task_kwargs=dict(primitive_fingertip_collisions=True))
    assert env.flat.meta["disable_hand_collisions"] is False
    obs, rew, dis, st = env.reset()
    asp = env.action_spec()
    a = torch.tensor(np.tile(0.5 * (asp.minimum + asp.maximum), (4, 1)), dtype=torch.float64, device="cuda")
    for _ in range(5):
        obs, rew, dis, st = env.step(a)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert (env.batch.get_diag()["flags"] & 3 == 0).all()
    env.close()


def test_beyond_the_record_capacity_the_deepest_contacts_are_kept():
    """Beyond its contact-record capacity (21; a cross-hand contact takes two records, every other contact one) the kernel keeps
    the deepest records (ordered by distance, a cross-hand contact as a unit), drops the shallow ones and flags the environment
    (diag bit 0) — deterministically; the oracle emulates the same rule (orc_set_record_cap), so the trajectories still agree.
    Hands flailing over their whole joint ranges need up to ~17 records: the capacity is lowered to 12 for the test."""
    CAP = 12
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2, scene_kw=HC)
    batch.debug_set_record_cap(CAP)
    om = O.OracleModel(fm); om.set_option(maxcon=60); om.set_record_cap(CAP)
    od = O.OracleData(om); od.reset()
    nv, nu = fm.meta["nv"], fm.meta["nu"]
    batch.set_state(np.zeros((2, nv)), np.zeros((2, nv)), np.zeros((2, nv)))
    od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
    names = fm.meta["actuator_names"]
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    rng = np.random.default_rng(3)
    target = rng.random(nu)
    worst = 0.0
    overflowed = 0
    for t in range(700):
        if t % 100 == 0:
            target = rng.random(nu)
        if t % 10 == 0:
            ctrl = lo + (hi - lo) * target  # full-range targets
            for i, n in enumerate(names):
                if n.endswith("forearm_tx"):
                    ctrl[i] = (-0.2 if "rh_" in n else 0.2) * min(1.0, t / 200.0)
                if n.endswith("forearm_ty"):
                    ctrl[i] = 0.0
            ct = _dev(np.tile(ctrl, (2, 1)))
        if t % 25 == 0 and t:
            batch.set_state(np.tile(od.qpos, (2, 1)), np.tile(od.qvel, (2, 1)), np.tile(od.qacc_warmstart, (2, 1)))
        batch.physics_step(ct.data_ptr(), 1)
        od.mj_step(ctrl)
        q, v, _, _ = batch.get_state()
        worst = max(worst, np.abs(q[0] - od.qpos).max())
        dg = batch.get_diag()
        over_dev, over_ref = bool(dg["flags"][0] & 1), bool(od.con_overflow)
        assert over_dev == over_ref, (t, dg["flags"][0], od.con_overflow, dg["ncon"][0], od.ncon)
        assert dg["ncon"][0] == od.ncon and dg["niter"][0] == od.solver_niter, (t, dg["ncon"][0], od.ncon, dg["niter"][0], od.solver_niter)
        overflowed += over_dev
        assert (q[0] == q[1]).all()
    assert overflowed > 10, overflowed
    assert worst < 1e-9, worst
    batch.close()

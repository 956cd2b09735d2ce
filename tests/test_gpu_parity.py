"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs,
plus size-independent properties at the full 4096-env batch.  Tolerances: state within 1e-9 absolute
per substep over 1000 substeps (north_star asks 1e-5 relative); rewards / observations 1e-9."""
import numpy as np
import pytest

import helpers
from oracle import oracle as O
from oracle import task_ref

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def test_native_library_is_loaded():
    from robopianist_b200 import native
    assert native.LIB_PATH.exists()
    native.lib()


def test_self_actuated_piano_trajectory():
    fm, model, batch = helpers.make_native("self_actuated_piano", nenv=3, lookahead=0, gravity_compensation=False)
    od = O.OracleData(O.OracleModel(fm))
    od.reset()
    rng = np.random.default_rng(1)
    nv, nu = fm.meta["nv"], fm.meta["nu"]
    batch.set_state(np.zeros((3, nv)), np.zeros((3, nv)), np.zeros((3, nv)))
    od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
    worst = 0.0
    for t in range(400):
        if t % 10 == 0:
            ctrl = rng.uniform(0, 1, nu) * fm.actuator_ctrlrange[:, 1] * (rng.random(nu) < 0.3)
            ct = _dev(np.tile(ctrl, (3, 1)))
        batch.physics_step(ct.data_ptr(), 1)
        od.mj_step(ctrl)
        q, v, _, _ = batch.get_state()
        worst = max(worst, np.abs(q[0] - od.qpos).max(), np.abs(v[0] - od.qvel).max() * 1e-2)
        assert (q[0] == q[2]).all()
    assert worst < 1e-11
    batch.close()


def test_hands_1000_substeps_match_oracle():
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2)
    om = O.OracleModel(fm); om.set_option(maxcon=24)
    od = O.OracleData(om); od.reset()
    nv = fm.meta["nv"]
    acts = helpers.twinkle_actions()
    batch.set_state(np.zeros((2, nv)), np.zeros((2, nv)), np.zeros((2, nv)))
    od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
    worst_q = worst_rel = 0.0
    contacts_seen = 0
    for t in range(1000):
        ctrl = acts[t // 10, :44]
        if t % 10 == 0:
            ct = _dev(np.tile(ctrl, (2, 1)))
        batch.physics_step(ct.data_ptr(), 1)
        od.mj_step(ctrl)
        q, v, _, _ = batch.get_state()
        worst_q = max(worst_q, np.abs(q[0] - od.qpos).max())
        worst_rel = max(worst_rel, (np.abs(v[0] - od.qvel) / (1e-3 + np.abs(od.qvel))).max())
        contacts_seen += od.ncon
        if t % 100 == 99:
            dg = batch.get_diag()
            assert dg["ncon"][0] == od.ncon and dg["niter"][0] == od.solver_niter and dg["flags"][0] == 0
    assert contacts_seen > 500      # the rollout really exercises contacts
    assert worst_q < 1e-9 and worst_rel < 1e-5
    batch.close()


@pytest.mark.parametrize("song,lookahead", [("TwinkleTwinkleRousseau", 1), ("CMajorScaleTwoHands", 3)])
def test_task_layer_matches_reference_logic(song, lookahead):
    """Full rp_step (ctrl write, 10 substeps, key state, 5 reward terms, obs, termination, auto-reset)."""
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", song=song, nenv=2, lookahead=lookahead)
    goal, fing = helpers.song_tables(song)
    ref = task_ref.RefEnv(fm, goal, fing, n_steps_lookahead=lookahead)
    N = 2
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda"); terms = torch.empty((N, 5), dtype=torch.float64, device="cuda")
    batch.reset(None, obs.data_ptr())
    rst, _, _, robs = ref.reset()
    torch.cuda.synchronize()
    assert np.abs(obs[0].cpu().numpy() - robs).max() < 1e-12
    acts = helpers.twinkle_actions()
    T = goal.shape[0]
    for t in range(T + 3):
        a = acts[t % 158]
        batch.step(_dev(np.tile(a, (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr(), terms.data_ptr())
        rst, rrew, rdis, robs = ref.step(a)
        torch.cuda.synchronize()
        assert int(st[0]) == rst, (t, int(st[0]), rst)
        assert np.abs(obs[0].cpu().numpy() - robs).max() < 1e-9, t
        if rst != task_ref.FIRST:
            assert abs(rew[0].item() - rrew) < 1e-9 and dis[0].item() == rdis, (t, rew[0].item(), rrew)
            tr = list(ref.reward_terms.values())
            assert np.abs(terms[0].cpu().numpy()[:len(tr)] - np.array(tr)).max() < 1e-9
    batch.close()


def test_full_batch_properties():
    """4096 envs: identical inputs -> bitwise identical outputs across envs and across runs; noisy envs diverge;
    episodes end after exactly T steps and auto-reset."""
    N = 4096
    acts = helpers.twinkle_actions()
    finals = []
    for run in range(2):
        fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=N)
        obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
        rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
        st = torch.empty(N, dtype=torch.uint8, device="cuda")
        batch.reset(None, obs.data_ptr())
        rng = np.random.default_rng(7)
        for t in range(12):
            a = np.tile(acts[t], (N, 1))
            a[N // 2:, :44] += rng.normal(0, 0.02, (N - N // 2, 44))
            batch.step(_dev(a).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        q, v, _, tix = batch.get_state()
        finals.append(q.copy())
        assert (q[: N // 2] == q[0]).all() and (tix == 12).all()
        assert np.abs(q[N // 2:] - q[0]).max() > 1e-6
        assert np.isfinite(q).all() and (batch.get_diag()["flags"] == 0).all()
        batch.close()
    assert (finals[0] == finals[1]).all()


def test_e2e_host_call_and_launch_count():
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=64)
    acts = helpers.twinkle_actions()
    batch.reset()
    n0 = batch.launches
    obs, rew, dis, st, terms = batch.step_host(np.tile(acts[0], (64, 1)))
    assert batch.launches == n0 + 1 and obs.shape == (64, 329) and np.isfinite(rew).all() and (st == 1).all()
    assert np.allclose(terms.sum(axis=1), rew)
    batch.close()


def test_golden_oracle_fixture_through_cabi():
    """CUDA path vs the committed oracle trajectory (tests/golden/oracle_twinkle_traj.npz; generator
    scripts/gen_oracle_golden.py): no oracle code runs here."""
    gold = np.load(helpers.GOLDEN / "oracle_twinkle_traj.npz")
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=1)
    nv = fm.meta["nv"]
    acts = helpers.twinkle_actions()
    batch.set_state(np.zeros((1, nv)), np.zeros((1, nv)), np.zeros((1, nv)))
    for t in range(60):
        batch.physics_step(_dev(acts[t:t + 1, :44]).data_ptr(), 10)
        if t in (9, 29, 59):
            q, _, _, _ = batch.get_state()
            assert np.abs(q[0] - gold[f"qpos_{t}"]).max() < 1e-9, t
    batch.close()


def test_heterogeneous_envs_match_their_own_oracles():
    """Every env of a batch gets different (seeded) action noise; each is compared with its own oracle
    instance: exercises per-env contact sets, iteration counts and early exits inside one launch."""
    E, steps = 6, 25
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=E)
    om = O.OracleModel(fm); om.set_option(maxcon=24)
    ods = [O.OracleData(om) for _ in range(E)]
    nv = fm.meta["nv"]
    acts = helpers.twinkle_actions()
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    rng = np.random.default_rng(12345)
    batch.set_state(np.zeros((E, nv)), np.zeros((E, nv)), np.zeros((E, nv)))
    for od in ods:
        od.reset()
    niters = set()
    for t in range(steps):
        ctrl = np.clip(acts[t, :44] + rng.normal(0, 0.05, (E, 44)) * (hi - lo) / 2, lo, hi)
        ctrl[0] = acts[t, :44]
        batch.physics_step(_dev(ctrl).data_ptr(), 10)
        q, v, _, _ = batch.get_state()
        dg = batch.get_diag()
        for e, od in enumerate(ods):
            od.control_step(ctrl[e])
            assert np.abs(q[e] - od.qpos).max() < 1e-9, (t, e)
            # (od.ncon after control_step belongs to the post-integration mj_step1, not to the last solve)
            assert dg["niter"][e] == od.solver_niter, (t, e)
            niters.add(int(od.solver_niter))
    assert len(niters) > 3  # the envs really ran different sweep counts
    batch.close()


def test_key_shared_by_both_hands_matches_oracle():
    """Left hand shifted over the right hand's keys (forearm_tx + 0.22): from control step 23 on key 58
    carries contacts of both hands, which couples the two per-hand Gauss-Seidel streams (the kernel's
    shared_key path: rows stepped in global order with a block barrier per row)."""
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2)
    om = O.OracleModel(fm); om.set_option(maxcon=24)
    od = O.OracleData(om); od.reset()
    nv = fm.meta["nv"]
    gh = fm.arrays["geom_hand"]
    acts = helpers.twinkle_actions()
    batch.set_state(np.zeros((2, nv)), np.zeros((2, nv)), np.zeros((2, nv)))
    shared_substeps = 0
    for t in range(36):
        ctrl = acts[t, :44].copy()
        ctrl[42] += 0.22
        ct = _dev(np.tile(ctrl, (2, 1)))
        for s in range(10):
            batch.physics_step(ct.data_ptr(), 1)
            od.mj_step(ctrl)
            keys = {}
            for c in od.contacts():
                if c["key"] >= 0:
                    keys.setdefault(c["key"], set()).add(int(gh[c["geom1"]]))
            shared_substeps += any(len(v) > 1 for v in keys.values())
        q, v, _, _ = batch.get_state()
        dg = batch.get_diag()
        assert np.abs(q[0] - od.qpos).max() < 1e-9 and (q[0] == q[1]).all(), t
        assert dg["ncon"][0] == od.ncon and dg["niter"][0] == od.solver_niter and dg["flags"][0] == 0, t
    assert shared_substeps > 50
    batch.close()


@pytest.mark.parametrize("iterations", [10, 20, 100])
def test_pgs_iteration_cap_sweep(iterations):
    """BASELINE config 5 in miniature: the iteration cap is honoured identically on both sides."""
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=1)
    batch.set_option(iterations=iterations)
    om = O.OracleModel(fm); om.set_option(iterations=iterations, maxcon=24)
    od = O.OracleData(om); od.reset()
    nv = fm.meta["nv"]
    acts = helpers.twinkle_actions()
    batch.set_state(np.zeros((1, nv)), np.zeros((1, nv)), np.zeros((1, nv)))
    capped = 0
    for t in range(30):
        batch.physics_step(_dev(acts[t:t + 1, :44]).data_ptr(), 10)
        od.control_step(acts[t, :44])
        q, _, _, _ = batch.get_state()
        dg = batch.get_diag()
        assert np.abs(q[0] - od.qpos).max() < 1e-9, t
        assert dg["niter"][0] == od.solver_niter <= iterations
        capped += od.solver_niter == iterations
    assert iterations == 100 or capped > 0
    batch.close()


def test_musical_metrics_match_the_wrapper_logic():
    """Device-side precision / recall / F1 (rp_get_metrics) vs the restated MidiEvaluationWrapper
    (oracle/task_ref.py, pinned to sklearn in tests/test_oracle.py) over one whole episode."""
    song = "TwinkleTwinkleRousseau"
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", song=song, nenv=2)
    goal, fing = helpers.song_tables(song)
    ref = task_ref.RefEnv(fm, goal, fing, n_steps_lookahead=1)
    N = 2
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda")
    batch.reset(None, obs.data_ptr()); ref.reset()
    acts = helpers.twinkle_actions()
    for t in range(goal.shape[0]):
        a = acts[t % 158]
        batch.step(_dev(np.tile(a, (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        rst, _, _, _ = ref.step(a)
    torch.cuda.synchronize()
    assert rst == task_ref.LAST and int(st[0]) == task_ref.LAST
    m = batch.get_metrics()
    assert (m["episodes"] == 1).all() and (m["running_steps"] == goal.shape[0]).all()
    for k, v in ref.musical_metrics.items():
        assert abs(m[k][0] - v) < 1e-12 and m[k][0] == m[k][1], (k, m[k], v)
    assert 0.3 < m["recall"][0] <= 1.0   # the recorded policy really presses goal keys
    # the next step auto-resets: running sums restart, the last episode's means stay
    batch.step(_dev(np.tile(acts[0], (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    m2 = batch.get_metrics()
    assert (m2["running_steps"] == 0).all() and m2["f1"][0] == m["f1"][0]
    batch.close()


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["piano_with_shadow_hands", "self_actuated_piano"])
def test_midi_event_stream_matches_midi_module(task):
    """Device-side MIDI messages (rp_midi_enable / rp_get_midi) vs the restated MidiModule.after_substep
    (models/piano/midi_module.py:47-98) over the oracle's substeps: same messages, same order, same times;
    the list restarts with the next episode."""
    song = "TwinkleTwinkleRousseau"
    fm, model, batch = helpers.make_native(task, song=song, nenv=3)
    goal, fing = helpers.song_tables(song)
    hands = task == "piano_with_shadow_hands"
    ref = task_ref.RefEnv(fm, goal, fing, kind=0 if hands else 1, n_steps_lookahead=1)
    ref.record_midi = True
    N = 3
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda")
    batch.midi_enable(16384)
    batch.reset(None, obs.data_ptr()); ref.reset()
    if hands:
        acts = helpers.twinkle_actions()
        acts[:, -1] = (np.arange(len(acts)) // 7) % 2   # exercise the pedal as well
    else:
        rng = np.random.default_rng(3)
        qmax = fm.actuator_ctrlrange[:, 1]
        acts = np.concatenate([rng.uniform(0, 1, (goal.shape[0], 88)) ** 6 * qmax, rng.integers(0, 2, (goal.shape[0], 1))], axis=1)
    T = goal.shape[0]
    for t in range(T):
        a = acts[t % len(acts)]
        batch.step(_dev(np.tile(a, (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        ref.step(a)
        if t in (0, 40, T - 1):
            got = batch.get_midi(fm.meta["timestep"])
            assert got[0] == got[2]
            assert len(got[0]) == len(ref.midi_messages), (t, len(got[0]), len(ref.midi_messages))
            for g, r in zip(got[0], ref.midi_messages):
                assert g[1:] == r[1:] and abs(g[0] - r[0]) < 1e-12, (t, g, r)
    kinds = {m[1] for m in ref.midi_messages}
    assert {"note_on", "note_off", "sustain_on", "sustain_off"} <= kinds and len(ref.midi_messages) > 20
    # stepping past the end of the episode starts an empty list
    batch.step(_dev(np.tile(acts[0] * 0, (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    assert all(len(m) == 0 for m in batch.get_midi(fm.meta["timestep"]))
    batch.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["substeps", "heterogeneous", "shared_key", "task", "midi"])
def test_phase_locked_blocks_match_the_oracle(monkeypatch, case):
    """The phase-locked launch shape (rp_set_block_envs(RP_BLOCK_ENVS): several environments per thread block,
    named barriers per environment, a block-wide barrier per substep) against the same oracles as the
    one-environment-per-block shape small batches get by default."""
    from robopianist_b200 import native
    monkeypatch.setenv("ROBOPLAY_BLOCK_ENVS", str(native.BLOCK_ENVS))
    if case == "substeps":
        test_hands_1000_substeps_match_oracle()
    elif case == "heterogeneous":
        test_heterogeneous_envs_match_their_own_oracles()
    elif case == "shared_key":
        test_key_shared_by_both_hands_matches_oracle()
    elif case == "task":
        test_task_layer_matches_reference_logic("TwinkleTwinkleRousseau", 1)
    else:
        test_midi_event_stream_matches_midi_module("piano_with_shadow_hands")


@pytest.mark.gpu
def test_launch_shapes_are_bitwise_identical():
    """13 environments (multi-environment blocks, the last one partly empty) stepped through an episode end
    (auto-reset of some environments only: wrong-press termination), a masked reset and plain physics steps:
    every output of the phase-locked shape equals the 1-per-block shape bit for bit."""
    from robopianist_b200 import native
    N = 13
    outs = []
    for be in (1, native.BLOCK_ENVS):
        fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=N, wrong_press_termination=True)
        batch.set_block_envs(be)
        lo = np.concatenate([fm.actuator_ctrlrange[:, 0], [0.0]]); hi = np.concatenate([fm.actuator_ctrlrange[:, 1], [1.0]])
        rng = np.random.default_rng(5)
        acts = helpers.twinkle_actions()
        obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
        rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
        st = torch.empty(N, dtype=torch.uint8, device="cuda")
        batch.reset(None, obs.data_ptr())
        rec = []
        for t in range(30):
            a = np.clip(acts[t] + rng.normal(0, 0.1, (N, 45)) * (hi - lo) / 2, lo, hi)
            batch.step(_dev(a).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
            rec += [obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), st.cpu().numpy().copy()]
            if t == 12:
                mask = torch.tensor([i % 3 == 0 for i in range(N)], dtype=torch.uint8, device="cuda")
                batch.reset(mask.data_ptr(), obs.data_ptr())
                rec.append(obs.cpu().numpy().copy())
        batch.physics_step(_dev(a[:, :44]).data_ptr(), 3)
        q, v, w, ti = batch.get_state()
        rec += [q, v, w, ti]
        outs.append(rec)
        batch.close()
    assert any(x.dtype == np.uint8 and (x == 2).any() for x in outs[0])  # some environments ended their episode early
    for x, y in zip(*outs):
        assert np.array_equal(x, y)


@pytest.mark.gpu
def test_uncovered_collider_proximity_flag():
    """diag flag 8: a hand collider outside the narrowphase's coverage (palm boxes, ...) within reach of the piano.
    Never on the recorded replay; set when the state is forced far outside the joint limits (forearm 15 cm in, wrist
    bent 90 degrees: tests/test_oracle.py restates the geometry)."""
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2)
    N = 2
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda")
    acts = helpers.twinkle_actions()
    batch.reset(None, obs.data_ptr())
    for t in range(60):
        a = torch.tensor(np.tile(acts[t], (N, 1)), dtype=torch.float64, device="cuda")
        batch.step(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        assert not (batch.get_diag()["flags"] & 8).any()
    q, v, w, tidx = batch.get_state()
    q[1, 89] = 0.15; q[1, 91] = 1.57
    batch.set_state(q, v, w, tidx)
    a = torch.tensor(np.tile(acts[60], (N, 1)), dtype=torch.float64, device="cuda")
    batch.step(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    f = batch.get_diag()["flags"]
    assert not (f[0] & 8) and (f[1] & 8)
    batch.close()


@pytest.mark.parametrize("reward_type", [0, 1])
def test_self_actuated_reward_identity_on_the_cuda_path(reward_type):
    """self_actuated_piano_test.py:139-166 on the CUDA path: random controls in the action spec; at every step the
    returned reward equals RewardType.get()(concat(activation, sustain_activation), goal_current) — negative L2
    distance (self_actuated_piano.py:49-57) or negative binary cross entropy (:38-47) — recomputed here in numpy from
    the activation the kernel reports and the goal row of the previous observation; observation, step type and
    discount against the restated task (oracle/task_ref.py) through the end of the episode and the auto-reset."""
    song = "CMajorScaleTwoHands"
    fm, model, batch = helpers.make_native("self_actuated_piano", song=song, nenv=3, lookahead=1, gravity_compensation=False,
                                           reward_type=reward_type)
    goal, fing = helpers.song_tables(song)
    ref = task_ref.RefEnv(fm, goal, fing, kind=1, n_steps_lookahead=1, reward_type=reward_type)
    N = 3
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda")
    batch.reset(None, obs.data_ptr()); ref.reset()
    torch.cuda.synchronize()
    lo = np.append(fm.actuator_ctrlrange[:, 0], 0.0); hi = np.append(fm.actuator_ctrlrange[:, 1], 1.0)
    rng = np.random.default_rng(12345)
    prev_goal = obs[0].cpu().numpy()[89:89 + 89].copy()      # obs = activation(88), sustain(1), goal((L+1)*89)
    nonzero = 0
    for t in range(goal.shape[0] + 2):
        a = rng.uniform(lo, hi)
        batch.step(_dev(np.tile(a, (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        rst, rrew, rdis, robs = ref.step(a)
        torch.cuda.synchronize()
        o = obs[0].cpu().numpy()
        assert int(st[0]) == rst and np.abs(o - robs).max() < 1e-12
        assert torch.equal(obs[0], obs[2]) and torch.equal(rew[0:1], rew[2:3])
        if rst != task_ref.FIRST:
            pred = np.append(batch.get_aux()["activation"][0].astype(np.float64), float(a[-1] >= 0.5))
            assert np.array_equal(pred, o[:89])              # the observables are the activations themselves
            if reward_type == 0:
                want = -np.sqrt(np.sum((pred - prev_goal) ** 2))
            else:
                want = np.sum(prev_goal * np.log(pred + 1e-6) + (1 - prev_goal) * np.log(1 - pred + 1e-6))
            assert abs(float(rew[0]) - want) < 1e-12 and abs(float(rew[0]) - rrew) < 1e-12 and float(dis[0]) == rdis
            nonzero += want != 0
        prev_goal = o[89:89 + 89].copy()
    assert nonzero > 50
    batch.close()


def test_diverged_environment_ends_its_episode_and_auto_resets():
    """Non-finite state (mjWARN_BADQPOS territory; dm_control raises PhysicsError): the environment is flagged (diag
    bit 1), returns LAST with discount 0 and reward 0, and the next step starts a fresh episode; its neighbours are
    bit-for-bit what they are without the incident."""
    acts = helpers.twinkle_actions()
    outs = []
    for poison in (False, True):
        fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=3)
        N = 3
        obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
        rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
        st = torch.empty(N, dtype=torch.uint8, device="cuda")
        batch.reset(None, obs.data_ptr())
        rec = []
        for t in range(8):
            if poison and t == 4:
                q, v, w, tidx = batch.get_state()
                v[1, 95] = np.inf
                batch.set_state(q, v, w, tidx)
            batch.step(_dev(np.tile(acts[t], (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
            torch.cuda.synchronize()
            rec.append((obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), dis.cpu().numpy().copy(), st.cpu().numpy().copy(),
                        batch.get_diag()["flags"].copy()))
        outs.append(rec)
        batch.close()
    clean, bad = outs
    o, r, d, s_, f = bad[4]
    assert (f[1] & 2) and s_[1] == 2 and d[1] == 0.0 and r[1] == 0.0 and not (f[0] & 2) and not (f[2] & 2)
    o, r, d, s_, f = bad[5]
    assert s_[1] == 0 and r[1] == 0.0 and d[1] == 1.0 and np.isfinite(o[1]).all() and not (f[1] & 2)
    assert bad[6][3][1] == 1 and np.isfinite(bad[6][0][1]).all()
    for t in range(8):
        for k in range(4):
            assert np.array_equal(clean[t][k][[0, 2]], bad[t][k][[0, 2]])

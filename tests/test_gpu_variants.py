"""GPU parity for the model-changing task kwargs (reduced_action_space, attachment_yaw, forearm_dofs subsets,
randomize_hand_positions): CUDA path through the C-ABI against the CPU oracle, which keeps jointless bodies as
separate bodies (the kernel fuses them into their parents' links on the host) — state to 1e-9 per substep, equal
contact and sweep counts; the full task layer against the restated reference task."""
import copy

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from oracle import task_ref
from robopianist_b200.scene import cache

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

VARIANTS = {
    "ras": dict(reduced_action_space=True),
    "yaw10": dict(attachment_yaw=10.0),
    "fdx": dict(forearm_dofs=("forearm_tx",)),
    "ras_fd": dict(forearm_dofs=(), reduced_action_space=True),
}


def _dev(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_substeps_match_oracle(name):
    full = cache.load_scene(gravity_compensation=True)
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2, scene_kw=VARIANTS[name])
    om = O.OracleModel(fm); om.set_option(maxcon=24)
    od = O.OracleData(om); od.reset()
    nv, nu = fm.meta["nv"], fm.meta["nu"]
    acts = helpers.reduced_actions(full, fm, helpers.twinkle_actions())
    assert acts.shape[1] == nu + 1 == model.act_dim
    batch.set_state(np.zeros((2, nv)), np.zeros((2, nv)), np.zeros((2, nv)))
    od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
    worst_q = worst_rel = 0.0
    contacts_seen = 0
    for t in range(600):
        ctrl = acts[t // 10, :nu]
        if t % 10 == 0:
            ct = _dev(np.tile(ctrl, (2, 1)))
        batch.physics_step(ct.data_ptr(), 1)
        od.mj_step(ctrl)
        q, v, _, _ = batch.get_state()
        assert q.shape == (2, nv)
        worst_q = max(worst_q, np.abs(q[0] - od.qpos).max())
        worst_rel = max(worst_rel, (np.abs(v[0] - od.qvel) / (1e-3 + np.abs(od.qvel))).max())
        contacts_seen += od.ncon
        if t % 50 == 49:
            dg = batch.get_diag()
            assert dg["ncon"][0] == od.ncon and dg["niter"][0] == od.solver_niter and (dg["flags"][0] & 7) == 0, (t, dg, od.ncon, od.solver_niter)
    assert worst_q < 1e-9 and worst_rel < 1e-5, (worst_q, worst_rel)
    if name != "ras_fd":  # without forearm dofs the hands cannot follow the recorded replay down to the keys
        assert contacts_seen > 100
    batch.close()


def test_reduced_action_space_task_layer():
    """Observation (23 joints per hand), the reward terms and termination of the reduced-action-space model against the
    restated reference task, over an episode and through the auto-reset."""
    full = cache.load_scene(gravity_compensation=True)
    kw = dict(reduced_action_space=True)
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2, scene_kw=kw)
    goal, fing = helpers.song_tables()
    ref = task_ref.RefEnv(fm, goal, fing, n_steps_lookahead=1)
    N = 2
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda"); terms = torch.empty((N, 5), dtype=torch.float64, device="cuda")
    assert model.obs_dim == 88 + 1 + 46 + 2 * 89 + 10 and model.act_dim == 39
    batch.reset(None, obs.data_ptr())
    rst, _, _, robs = ref.reset()
    torch.cuda.synchronize()
    assert robs.shape == (model.obs_dim,) and np.abs(obs[0].cpu().numpy() - robs).max() < 1e-12
    acts = helpers.reduced_actions(full, fm, helpers.twinkle_actions())
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


def _shifted(fm, offset):
    """The flat model with both hands' attachment frames moved by `offset` along y (Entity.shift_pose)."""
    out = copy.deepcopy(fm)
    for b in out.arrays["hand_rootbody"]:
        out.arrays["body_pos"][b, 1] += offset
    return out


def test_randomize_hand_positions():
    """piano_with_shadow_hands.py:491-499: one U(-0.05, 0.05) shift along y per episode and environment, the same for
    both hands, accumulating over episodes; the physics of a shifted environment equals the oracle's on a model whose
    hand root bodies are moved by that offset."""
    N = 6
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=N)
    assert (batch.get_hand_offsets() == 0).all()
    batch.randomize_hand_positions(True, seed=123)
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda")
    batch.reset(None, obs.data_ptr())
    off1 = batch.get_hand_offsets()
    assert (np.abs(off1) <= 0.05).all() and len(np.unique(off1)) == N and np.abs(off1).max() > 1e-3
    acts = helpers.twinkle_actions()
    oracles = []
    for e in (0, 3):
        om = O.OracleModel(_shifted(fm, off1[e])); om.set_option(maxcon=24)
        od = O.OracleData(om); od.reset()
        oracles.append((e, om, od))
    for t in range(30):
        batch.step(_dev(np.tile(acts[t], (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        for e, om, od in oracles:
            od.control_step(acts[t, :44])
    q, v, _, _ = batch.get_state()
    for e, om, od in oracles:
        assert np.abs(q[e] - od.qpos).max() < 1e-9 and np.abs(v[e] - od.qvel).max() < 1e-6
    assert np.abs(q[0] - q[3]).max() > 1e-6  # different offsets, different trajectories
    assert (batch.get_hand_offsets() == off1).all()  # drawn at episode starts only
    # a masked reset draws again for the selected environments and accumulates
    mask = torch.zeros(N, dtype=torch.uint8, device="cuda"); mask[1] = 1; mask[4] = 1
    batch.reset(mask.data_ptr(), obs.data_ptr())
    off2 = batch.get_hand_offsets()
    keep = np.array([0, 2, 3, 5])
    assert (off2[keep] == off1[keep]).all() and (off2[[1, 4]] != off1[[1, 4]]).all()
    assert (np.abs(off2[[1, 4]] - off1[[1, 4]]) <= 0.05).all()
    # same seed, same draws; switching the feature off clears the offsets
    fm2, model2, batch2 = helpers.make_native("piano_with_shadow_hands", nenv=N)
    batch2.randomize_hand_positions(True, seed=123)
    batch2.reset(None, obs.data_ptr())
    assert (batch2.get_hand_offsets() == off1).all()
    batch2.randomize_hand_positions(False)
    assert (batch2.get_hand_offsets() == 0).all()
    # enable=2 (suite.load(recompile_physics=True)): every episode's shift starts from the unshifted pose
    batch2.randomize_hand_positions(2, seed=5)
    batch2.reset(None, obs.data_ptr()); o1 = batch2.get_hand_offsets()
    batch2.reset(None, obs.data_ptr()); o2 = batch2.get_hand_offsets()
    batch2.reset(None, obs.data_ptr()); o3 = batch2.get_hand_offsets()
    assert (np.abs(np.stack([o1, o2, o3])) <= 0.05).all() and (o1 != o2).all() and (o2 != o3).all()
    batch.close(); batch2.close()


def test_suite_load_with_the_reference_training_kwargs():
    """The task kwargs of the reference's RL training recipe that this path covers: reduced action space, gravity
    compensation, primitive fingertips, lookahead 10, trimmed silence (+ randomised hand positions)."""
    from robopianist_b200 import suite
    env = suite.load("RoboPianist-debug-TwinkleTwinkleRousseau-v0", seed=1, num_envs=8, task_kwargs=dict(
        reduced_action_space=True, gravity_compensation=True, primitive_fingertip_collisions=True, disable_hand_collisions=True,
        n_steps_lookahead=10, trim_silence=True, randomize_hand_positions=True, attachment_yaw=0.0))
    spec = env.observation_spec()
    assert spec["rh_shadow_hand/joints_pos"].shape == (23,) and spec["lh_shadow_hand/joints_pos"].shape == (23,)
    assert spec["goal"].shape == (11 * 89,)
    asp = env.action_spec()
    assert asp.shape == (39,) and asp.name.split("\t")[-1] == "sustain" and "A_THJ5" not in asp.name
    obs, rew, dis, st = env.reset()
    assert (st == 0).all() and len(np.unique(env.batch.get_hand_offsets())) == 8
    a = torch.tensor(np.tile(0.5 * (asp.minimum + asp.maximum), (8, 1)), dtype=torch.float64, device="cuda")
    for _ in range(3):
        obs, rew, dis, st = env.step(a)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and (st == 1).all()
    env.close()


def test_multi_device_env_matches_one_device():
    """`suite.load(..., devices=[0, 1])`: one process, the batch split over two GPUs — bitwise the same results as the whole
    batch on one (environments are independent; needs two devices)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from robopianist_b200 import suite
    kw = dict(primitive_fingertip_collisions=True, disable_hand_collisions=True, gravity_compensation=True, trim_silence=True)
    N = 64
    one = suite.load("RoboPianist-debug-TwinkleTwinkleRousseau-v0", num_envs=N, task_kwargs=kw)
    two = suite.load("RoboPianist-debug-TwinkleTwinkleRousseau-v0", num_envs=N, devices=[0, 1], task_kwargs=kw)
    assert [e.device.index for e in two.envs] == [0, 1] and two.ranges == [(0, 32), (32, 64)]
    acts = helpers.twinkle_actions()
    rng = np.random.default_rng(0)
    one.reset(); two.reset()
    for t in range(12):
        a = np.tile(acts[t], (N, 1)); a[:, :44] += rng.normal(0, 0.02, (N, 44))
        at = torch.tensor(a, dtype=torch.float64, device="cuda:0")
        o1, r1, d1, s1 = one.step(at)
        o2, r2, d2, s2 = two.step(at)
    torch.cuda.synchronize(0); two.synchronize()
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2) and torch.equal(s1, s2)
    one.close(); two.close()


def test_checkpoint_resumes_bit_for_bit():
    """rp_checkpoint_save / rp_checkpoint_load (BatchedPianoEnv.state_dict / load_state_dict): a batch restored in the middle of
    its episodes — random hand shifts, MIDI recording and hand collisions enabled, some environments past an auto-reset —
    continues exactly like the one that kept running; a blob of another shape is refused."""
    from robopianist_b200 import native, suite
    kw = dict(primitive_fingertip_collisions=True, gravity_compensation=True, trim_silence=True, randomize_hand_positions=True,
              wrong_press_termination=True)
    N = 16

    def make():
        e = suite.load("RoboPianist-debug-TwinkleTwinkleRousseau-v0", seed=3, num_envs=N, task_kwargs=kw)
        e.enable_midi(512)
        return e

    a, b = make(), make()
    acts = helpers.twinkle_actions()
    rng = np.random.default_rng(2)

    def action(t):
        x = np.tile(acts[t % 158], (N, 1)); x[:, :44] += rng.normal(0, 0.05, (N, 44))
        return torch.tensor(x, dtype=torch.float64, device="cuda")

    a.reset()
    for t in range(15):
        a.step(action(t))
    sd = a.state_dict()
    tail = [action(t) for t in range(15, 30)]
    outs_a = []
    for x in tail:
        o, r, d, s = a.step(x)
        outs_a.append((o.clone(), r.clone(), d.clone(), s.clone()))
    b.reset()                      # different random shifts and state until the checkpoint is loaded
    b.load_state_dict(sd)
    for x, (o1, r1, d1, s1) in zip(tail, outs_a):
        o, r, d, s = b.step(x)
        assert torch.equal(o, o1) and torch.equal(r, r1) and torch.equal(d, d1) and torch.equal(s, s1)
    qa, qb = a.get_state(), b.get_state()
    assert all(np.array_equal(x, y) for x, y in zip(qa, qb))
    assert (a.batch.get_hand_offsets() == b.batch.get_hand_offsets()).all()
    ma, mb = a.get_all_midi_messages(), b.get_all_midi_messages()
    assert ma == mb
    other = suite.load("RoboPianist-debug-TwinkleTwinkleRousseau-v0", seed=3, num_envs=N + 1, task_kwargs=kw)
    with pytest.raises(native.NativeError):
        other.batch.restore(sd["device"])
    for e in (a, b, other):
        e.close()

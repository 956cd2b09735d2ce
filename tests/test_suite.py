"""Suite surface: mirrors the reference's structural tests
(robopianist/suite/suite_test.py:27-56, suite/tasks/piano_with_shadow_hands_test.py:74-272) against
the CUDA-backed environment.  CPU part: registry, kwargs and loud failures; GPU part: the dm_env
behaviour itself."""
import numpy as np
import pytest

from robopianist_b200 import native

from robopianist_b200 import suite
from robopianist_b200.music import trajectory
from robopianist_b200.music.sequence import NoteSeq

KW = dict(primitive_fingertip_collisions=True, disable_hand_collisions=True)


def _test_midi(dt=0.01):
    """C6 for 2 dt (right index), then G5 for 1 dt (left thumb) — piano_with_shadow_hands_test.py:28-50."""
    seq = NoteSeq()
    seq.add(84, 0.0, 2 * dt, 80, 1)
    seq.add(79, 2 * dt, 3 * dt, 80, 5)
    seq.total_time = 3 * dt
    return seq


def _env(control_timestep=0.01, **kw):
    from robopianist_b200.suite.env import BatchedPianoEnv, PianoEnv
    kw = dict(KW, control_timestep=control_timestep, n_steps_lookahead=0, **kw)
    return PianoEnv(BatchedPianoEnv(_test_midi(control_timestep), num_envs=1, **kw))


# ------------------------------------------------------------------------------------ CPU
def test_registry_and_unknown_name():
    assert "RoboPianist-debug-TwinkleTwinkleRousseau-v0" in suite.DEBUG and set(suite.DEBUG) <= set(suite.ALL)
    with pytest.raises(ValueError):
        suite.load("RoboPianist-debug-NoSuchSong-v0")


def test_unsupported_kwargs_fail_loudly():
    name = suite.DEBUG[0]
    with pytest.raises(NotImplementedError):
        suite.load(name)  # reference default = mesh fingertips: not implemented, must not be silently replaced
    try:  # hand collisions enabled (the reference default) are accepted: past the kwargs check the only obstacle is the device
        suite.load(name, task_kwargs=dict(primitive_fingertip_collisions=True)).close()
    except native.NativeError as e:
        assert "CUDA" in str(e)
    with pytest.raises(TypeError):
        suite.load(name, task_kwargs=dict(KW, no_such_kwarg=1))
    with pytest.raises(NotImplementedError):
        suite.load(name, legacy_step=False, task_kwargs=KW)


def test_song_without_fingering_switches_to_the_ot_reward():
    """MidiFile.has_fingering (music/midi_file.py:252-261): `part` defaults to 0, so a song whose notes all carry
    part 0 has NO fingering — the task then disables the fingering reward, drops the 10-dim `fingering` observable
    and adds the optimal-transport fingering reward (piano_with_shadow_hands.py:109-111, 130-141, 443-449)."""
    from robopianist_b200.suite import env as E
    plain = NoteSeq()
    plain.add(60, 0.0, 0.5); plain.add(64, 0.5, 1.0)
    plain.total_time = 1.0
    assert not plain.has_fingering()
    cfg = E.derive_task_config(plain, True, dict(E._TASK_KWARGS))
    assert cfg["disable_fingering_reward"] and cfg["ot_fingering_reward"]
    assert (cfg["fingering"] >= 0).any()            # the table still lists finger 0 for every note: not the criterion
    layout = E.obs_layout(True, cfg["n_steps_lookahead"], cfg["disable_fingering_reward"])
    assert "fingering" not in layout and next(reversed(layout.values())).stop == 319
    # one non-zero part among several distinct ones: fingering present
    cfg = E.derive_task_config(_test_midi(), True, dict(E._TASK_KWARGS))
    assert not cfg["disable_fingering_reward"] and not cfg["ot_fingering_reward"]
    assert next(reversed(E.obs_layout(True, 1, False).values())).stop == 329
    # all notes on the same non-zero part: a single distinct part is "no fingering" as well
    mono = NoteSeq(); mono.add(60, 0.0, 0.5, 80, 3); mono.add(62, 0.5, 1.0, 80, 3); mono.total_time = 1.0
    assert E.derive_task_config(mono, True, dict(E._TASK_KWARGS))["disable_fingering_reward"]
    # the keys-only task never uses the OT term
    assert not E.derive_task_config(plain, False, dict(E._TASK_KWARGS))["ot_fingering_reward"]


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_plain_midi_reaches_the_ot_reward_through_load(tmp_path):
    """The same rule end to end: a song without fingering loaded through the suite runs the OT kernel (reward term
    named ot_fingering_reward, obs 319 wide)."""
    from robopianist_b200.suite.env import BatchedPianoEnv, PianoEnv
    plain = NoteSeq()
    plain.add(60, 0.0, 0.1); plain.add(64, 0.1, 0.2)
    plain.total_time = 0.2
    env = PianoEnv(BatchedPianoEnv(plain, num_envs=1, **KW))
    ts = env.reset()
    assert "fingering" not in ts.observation and sum(v.size for v in ts.observation.values()) == 319
    ts = env.step(np.zeros(45))
    terms = env.task.reward_fn.reward_terms
    assert "ot_fingering_reward" in terms and "fingering_reward" not in terms
    assert 0.0 <= terms["ot_fingering_reward"] <= 1.0
    assert abs(sum(terms.values()) - ts.reward) < 1e-12
    env.close()


@pytest.mark.gpu
def test_augmentations_change_the_song_at_every_reset():
    """task_kwargs['augmentations'] (suite/variations.py): a new song per reset, drawn from the env's seeded
    RandomState exactly as the reference's _maybe_change_midi does (piano_with_shadow_hands.py:151-165); the goal
    observation and the episode length follow the augmented song."""
    from robopianist_b200.suite import variations as V
    from robopianist_b200.suite.env import BatchedPianoEnv, PianoEnv
    midi = _test_midi(0.05)
    aug = [V.MidiPitchShift(prob=1.0, shift_range=12), V.MidiTemporalStretch(prob=1.0, stretch_range=0.5)]
    env = PianoEnv(BatchedPianoEnv(midi, num_envs=1, seed=7, control_timestep=0.05, n_steps_lookahead=0, augmentations=aug, **KW))
    rs = np.random.RandomState(7)
    seen = set()
    for ep in range(4):
        ts = env.reset()
        want = V.apply(aug, midi, rs)
        traj = trajectory.NoteTrajectory.from_seq(want, 0.05)
        assert len(env.task._notes) == len(traj.notes)
        assert np.array_equal(ts.observation["goal"], traj.goal_table()[0].astype(np.float64))
        seen.add((want.notes[0].pitch, len(traj.notes)))
        n = 0
        while not ts.last():
            ts = env.step(np.zeros(45)); n += 1
        assert n == len(traj.notes)
    assert len(seen) > 1
    env.close()


@pytest.mark.gpu
def test_masked_reset_reports_first_for_the_reset_envs():
    import torch
    from robopianist_b200.suite.env import BatchedPianoEnv
    env = BatchedPianoEnv(_test_midi(0.05), num_envs=4, control_timestep=0.05, **KW)
    env.reset()
    a = torch.zeros((4, 45), dtype=torch.float64, device="cuda")
    env.step(a)
    mask = torch.tensor([0, 1, 0, 1], dtype=torch.uint8, device="cuda")
    _, rew, dis, st = env.reset(mask)
    torch.cuda.synchronize()
    assert st.tolist() == [1, 0, 1, 0] and dis.tolist()[1] == 1.0 and rew.tolist()[1] == 0.0 and rew.tolist()[3] == 0.0
    env.close()


@pytest.mark.gpu
def test_all_debug_envs_load_and_step():
    """suite_test.py:27-56: every debug env loads, FIRST has no reward/discount, obs matches the spec."""
    rng = np.random.default_rng(12345)
    for name in suite.DEBUG:
        midi_has_fingering = name not in ()
        env = suite.load(name, seed=12345, task_kwargs=dict(KW))
        spec = env.action_spec()
        assert spec.shape == (45,) and np.isfinite(spec.minimum).all() and np.isfinite(spec.maximum).all()
        ts = env.reset()
        assert ts.first() and ts.reward is None and ts.discount is None
        ospec = env.observation_spec()
        for _ in range(10):
            ts = env.step(rng.uniform(spec.minimum, spec.maximum))
            assert list(ts.observation) == list(ospec)
            for k, v in ts.observation.items():
                assert v.shape == ospec[k].shape and np.isfinite(v).all()
            assert np.isfinite(ts.reward)
        env.close()


@pytest.mark.gpu
def test_observables_and_action_layout():
    env = _env()
    keys = list(env.observation_spec())
    assert keys == ["piano/state", "piano/sustain_state", "rh_shadow_hand/joints_pos", "lh_shadow_hand/joints_pos",
                    "goal", "fingering"]
    names = env.action_spec().name.split("\t")
    assert len(names) == 45 and names[-1] == "sustain"
    assert all(n.startswith("rh_shadow_hand/") for n in names[:22]) and all(n.startswith("lh_shadow_hand/") for n in names[22:44])
    env.close()


@pytest.mark.gpu
def test_termination_and_discount():
    env = _env()
    zero = np.zeros(45)
    env.reset()
    for _ in range(3):
        ts = env.step(zero)
        assert ts.mid() and ts.discount == 1.0
    ts = env.step(zero)
    assert ts.last() and ts.discount == 1.0
    assert env.step(zero).first()  # dm_env: the step after LAST resets
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ct,secs", [(0.01, 0), (0.01, 0.1), (0.05, 0.01), (0.05, 1)])
def test_n_seconds_lookahead(ct, secs):
    env = _env(control_timestep=ct, n_seconds_lookahead=secs)
    assert env._batched.n_steps_lookahead == int(np.ceil(secs / ct))
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("L", [0, 1, 2, 5])
def test_goal_and_fingering_observables(L):
    from robopianist_b200.suite.env import BatchedPianoEnv, PianoEnv
    env = PianoEnv(BatchedPianoEnv(_test_midi(), num_envs=1, control_timestep=0.01, n_steps_lookahead=L, **KW))
    traj = trajectory.NoteTrajectory.from_seq(_test_midi(), 0.01)
    notes, sustains = traj.notes, traj.sustains
    assert len(notes) == 4
    ts = env.reset()
    for i in range(len(notes)):
        expected = np.zeros((L + 1, 89))
        for j, t in enumerate(range(i, min(i + L + 1, len(notes)))):
            expected[j, [n.key for n in notes[t]]] = 1.0
            expected[j, -1] = sustains[t]
        np.testing.assert_array_equal(ts.observation["goal"], expected.ravel())
        fing = np.zeros((2, 5))
        idx = [n.fingering for n in notes[i]]
        fing[0, [f for f in idx if f < 5]] = 1.0
        fing[1, [f - 5 for f in idx if f >= 5]] = 1.0
        np.testing.assert_array_equal(ts.observation["fingering"], fing.ravel())
        ts = env.step(np.zeros(45))
    env.close()


@pytest.mark.gpu
def test_failure_termination():
    """piano_with_shadow_hands_test.py:228-242 pushes every key down with qfrc_applied; the C-ABI has no
    applied-force input, so the same end state (all keys at their upper limit) is set directly."""
    env = _env(wrong_press_termination=True)
    env.reset()
    b = env._batched
    q, v, w, _ = b.get_state()
    q[0, :88] = b.flat.key_qmax
    b.set_state(q, v, w)
    ts = env.step(np.zeros(45))
    assert ts.last() and ts.discount == 0.0
    env.close()


@pytest.mark.gpu
def test_reward_terms_and_batched_surface():
    import torch
    env = suite.load("RoboPianist-debug-CMajorScaleTwoHands-v0", task_kwargs=dict(KW), num_envs=8)
    obs, rew, dis, st = env.reset()
    assert obs.shape == (8, 329) and obs.is_cuda and (st == 0).all()
    a = torch.zeros((8, 45), dtype=torch.float64, device="cuda")
    obs, rew, dis, st = env.step(a)
    torch.cuda.synchronize()
    assert (st == 1).all() and torch.allclose(env.reward_terms.sum(dim=1), rew)
    with pytest.raises(ValueError):
        env.step(a.float())
    env.close()
    one = suite.load("RoboPianist-debug-CMajorScaleTwoHands-v0", task_kwargs=dict(KW, disable_fingering_reward=True))
    one.reset(); one.step(np.zeros(45))
    assert "fingering_reward" not in one.task.reward_fn.reward_terms and "key_press_reward" in one.task.reward_fn.reward_terms
    assert "fingering" not in one.observation_spec()
    one.close()

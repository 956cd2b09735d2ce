"""Shared builders for the tests (scene, song tables, oracle / native handles)."""
import os
from pathlib import Path

import numpy as np

from robopianist_b200.music import library, trajectory
from robopianist_b200.scene import cache

GOLDEN = Path(__file__).resolve().parent / "golden"


def song_tables(name="TwinkleTwinkleRousseau", trim_silence=True, dt=0.05):
    seq = library.load(name)
    if trim_silence:
        seq = seq.trim_silence()
    traj = trajectory.NoteTrajectory.from_seq(seq, dt)
    return traj.goal_table(), traj.fingering_table()


def twinkle_actions():
    return np.load(GOLDEN / "twinkle_ctrl.npz")["action"]


def make_native(task="piano_with_shadow_hands", song="TwinkleTwinkleRousseau", gravity_compensation=True,
                nenv=1, lookahead=1, scene_kw=None, **task_kw):
    from robopianist_b200 import native
    fm = cache.load_scene(task, gravity_compensation=gravity_compensation, **(scene_kw or {}))
    goal, fing = song_tables(song)
    kind = 0 if task == "piano_with_shadow_hands" else 1
    model = native.Model(fm, goal, fing, kind=kind, n_steps_lookahead=lookahead, **task_kw)
    batch = native.Batch(model, nenv)
    if os.environ.get("ROBOPLAY_BLOCK_ENVS"):  # launch shape under test (results must not depend on it)
        batch.set_block_envs(int(os.environ["ROBOPLAY_BLOCK_ENVS"]))
    return fm, model, batch


def reduced_actions(fm_full, fm_reduced, actions):
    """The recorded 45-dim action replay mapped onto a model with fewer actuators (reduced action space, fewer forearm
    dofs): columns matched by actuator name, sustain last."""
    full = list(fm_full.meta["actuator_names"])
    cols = [full.index(n) for n in fm_reduced.meta["actuator_names"]] + [len(full)]
    return np.ascontiguousarray(actions[:, cols])

"""Environment surface over the CUDA step kernel.

Mirrors, for the one path this repository replaces, what the reference hands back from
``robopianist.suite.load`` (suite/__init__.py:50-93): a ``dm_env.Environment``-shaped object whose
``reset()`` / ``step(action)`` return ``TimeStep(step_type, reward, discount, observation)``
(composer.Environment semantics), plus the batched form a learner on the same
GPU wants: ``BatchedPianoEnv.step(actions[N, 45]) -> (obs[N, 329], reward[N], discount[N],
step_type[N])`` as CUDA tensors, one kernel launch per control step.

``dm_env`` is not installable here, so ``TimeStep`` / ``StepType`` / ``BoundedArray`` are restated
with the same field names and meaning.  There is no CPU path: constructing an environment without
the native library or without a CUDA device raises.
"""
from __future__ import annotations

import collections
import enum
from typing import Any, Mapping, NamedTuple, Optional

import numpy as np

from .. import native
from ..music import trajectory
from ..scene import cache

_TASK_KWARGS = dict(
    n_steps_lookahead=1, n_seconds_lookahead=None, trim_silence=False, wrong_press_termination=False,
    initial_buffer_time=0.0, disable_fingering_reward=False, disable_forearm_reward=False,
    disable_colorization=False, disable_hand_collisions=False, augmentations=None,
    energy_penalty_coef=5e-3, randomize_hand_positions=False,
    # PianoTask kwargs (tasks/base.py:96-107)
    gravity_compensation=False, change_color_on_activation=False, primitive_fingertip_collisions=False,
    reduced_action_space=False, attachment_yaw=0.0, forearm_dofs=("forearm_tx", "forearm_ty"),
    physics_timestep=0.005, control_timestep=0.05,
)


class StepType(enum.IntEnum):
    FIRST = 0
    MID = 1
    LAST = 2


class TimeStep(NamedTuple):
    step_type: Any
    reward: Any
    discount: Any
    observation: Any

    def first(self) -> bool:
        return self.step_type == StepType.FIRST

    def mid(self) -> bool:
        return self.step_type == StepType.MID

    def last(self) -> bool:
        return self.step_type == StepType.LAST


class Array:
    def __init__(self, shape, dtype, name=None):
        self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name

    def validate(self, value):
        value = np.asarray(value)
        if value.shape != self.shape:
            raise ValueError(f"Expected shape {self.shape} but found {value.shape}")
        return value

    def __repr__(self):
        return f"Array(shape={self.shape}, dtype={self.dtype}, name={self.name!r})"


class BoundedArray(Array):
    def __init__(self, shape, dtype, minimum, maximum, name=None):
        super().__init__(shape, dtype, name)
        self.minimum = np.asarray(minimum, dtype=dtype)
        self.maximum = np.asarray(maximum, dtype=dtype)


def _unsupported(kw: Mapping[str, Any]):
    """Reference kwargs this path does not implement must fail loudly, not be ignored."""
    if not kw["primitive_fingertip_collisions"]:
        raise NotImplementedError(
            "mesh fingertip colliders need a convex-mesh narrowphase (SURVEY.md §8(f) row 1); pass "
            "task_kwargs=dict(primitive_fingertip_collisions=True) — the north_star configuration")
    bad = [d for d in kw["forearm_dofs"] if d not in ("forearm_tx", "forearm_ty")]
    if bad:  # shadow_hand.py:41-52 also lists rotational forearm dofs; only the two slides are restated
        raise NotImplementedError(f"forearm_dofs {bad} are not implemented (forearm_tx / forearm_ty are)")


def derive_task_config(midi, hands: bool, kw: Mapping[str, Any]):
    """What the task constructor derives from the song and its kwargs (piano_with_shadow_hands.py:98-141), as plain
    host data — no device needed: trimmed song, note trajectory, goal / fingering tables, lookahead, and which of
    the fingering rewards is active."""
    control_timestep = float(kw["control_timestep"])
    if kw["trim_silence"]:
        midi = midi.trim_silence()
    L = int(kw["n_steps_lookahead"])
    if kw["n_seconds_lookahead"] is not None:  # piano_with_shadow_hands.py:103-106
        L = int(np.ceil(kw["n_seconds_lookahead"] / control_timestep))
    traj = trajectory.NoteTrajectory.from_seq(midi, control_timestep)
    if kw["initial_buffer_time"]:
        traj = traj.add_initial_buffer_time(kw["initial_buffer_time"])
    # MidiFile.has_fingering (music/midi_file.py; sequence.py restates it): more than one distinct `part` among the
    # notes and at least one non-zero — `Note.part` defaults to 0, so a plain .mid has none.  Evaluated after
    # trim_silence, like the reference (piano_with_shadow_hands.py:100-111).  No fingering in the song disables the
    # fingering reward (and drops the `fingering` observable); whenever it is disabled the task adds the
    # optimal-transport fingering reward in its place (:130-141, :333-369).
    disable_fingering = bool(kw["disable_fingering_reward"]) or not midi.has_fingering()
    return dict(midi=midi, trajectory=traj, goal=traj.goal_table(), fingering=traj.fingering_table(),
                n_steps_lookahead=L, control_timestep=control_timestep, disable_fingering_reward=disable_fingering,
                ot_fingering_reward=bool(hands and disable_fingering))


def obs_layout(hands: bool, n_steps_lookahead: int, disable_fingering_reward: bool, hand_dofs: int = 26):
    """Observation name -> slice of the flat observation row (SURVEY.md App. A order).  hand_dofs: joints per hand
    (26 = 24 hand joints + 2 forearm slides; 23 with reduced_action_space, fewer with fewer forearm_dofs)."""
    L, nl = n_steps_lookahead, hand_dofs
    out, off = collections.OrderedDict(), 0

    def add(name, n):
        nonlocal off
        out[name] = slice(off, off + n)
        off += n

    if hands:
        add("piano/state", 88); add("piano/sustain_state", 1)
        add("rh_shadow_hand/joints_pos", nl); add("lh_shadow_hand/joints_pos", nl)
        add("goal", (L + 1) * 89)
        if not disable_fingering_reward:
            add("fingering", 10)
    else:
        add("piano/activation", 88); add("piano/sustain_activation", 1)
        add("goal", (L + 1) * 89)
    return out


class BatchedPianoEnv:
    """N independent PianoWithShadowHands (or SelfActuatedPiano) environments on one GPU.

    All tensors are float64 CUDA tensors owned by this object and overwritten by the next call.
    Episodes auto-reset with dm_env semantics: the step after LAST returns FIRST (reward 0,
    discount 1) and ignores its action.
    """

    def __init__(self, midi, num_envs: int = 1, device: int = 0, task: str = "piano_with_shadow_hands",
                 iterations: Optional[int] = None, tolerance: Optional[float] = None,
                 reward_type: int = 0, seed: Optional[int] = None, recompile_physics: bool = False, **task_kwargs):
        import torch

        unknown = set(task_kwargs) - set(_TASK_KWARGS)
        if unknown:
            raise TypeError(f"unexpected task kwargs {sorted(unknown)}")
        kw = dict(_TASK_KWARGS, **task_kwargs)
        hands = task == "piano_with_shadow_hands"
        if hands:
            _unsupported(kw)
        if not torch.cuda.is_available():
            raise native.NativeError("BatchedPianoEnv needs a CUDA device (there is no CPU path)")
        self.torch = torch
        self.task_name = task
        cfg = derive_task_config(midi, hands, kw)
        # augmentations (suite/variations.py): applied to the initial song at every full reset(); the whole batch gets the
        # same draw (one song table per batch on the device), episodes that auto-reset inside the kernel keep the
        # current song until the next reset() call.  The reference draws per episode of its single environment.
        self._augmentations = list(kw["augmentations"]) if kw["augmentations"] else None
        self._initial_midi = cfg["midi"]
        self._initial_buffer_time = float(kw["initial_buffer_time"])
        self.random_state = np.random.RandomState(seed)
        self.control_timestep = cfg["control_timestep"]
        midi = self.midi = cfg["midi"]
        traj = self.trajectory = cfg["trajectory"]
        self._notes, self._sustains = traj.notes, traj.sustains
        goal, fing = cfg["goal"], cfg["fingering"]
        self.goal_table, self.fingering_table = goal, fing
        L = self.n_steps_lookahead = cfg["n_steps_lookahead"]
        self.disable_fingering_reward = cfg["disable_fingering_reward"]
        self.disable_forearm_reward = bool(kw["disable_forearm_reward"])
        self.flat = cache.load_scene(task, gravity_compensation=bool(kw["gravity_compensation"]),
                                     physics_timestep=float(kw["physics_timestep"]),
                                     control_timestep=self.control_timestep,
                                     forearm_dofs=tuple(kw["forearm_dofs"]),
                                     reduced_action_space=bool(kw["reduced_action_space"]),
                                     disable_hand_collisions=bool(kw["disable_hand_collisions"]),
                                     attachment_yaw=float(kw["attachment_yaw"]))
        n_ign = int(self.flat.meta.get("n_ignored_colliders", 0))
        if hands and n_ign:
            import warnings
            warnings.warn(
                f"{n_ign} hand colliders of the reference model (palm / knuckle / wrist boxes and cylinders, forearm "
                "mesh) are outside the narrowphase's coverage and do not collide with the piano; environments in "
                "which one of them comes within reach of the piano are flagged (rp_get_diag flags bit 3, "
                "BatchedPianoEnv.uncovered_collider_flags())"
                + ("" if kw["disable_hand_collisions"] else "; with hand collisions enabled the sphere / capsule geoms of the fingers "
                   "and thumbs collide with each other (within a hand and across the hands), pairs that involve one of the "
                   "uncovered colliders do not, and the forearm reward tests the forearms' oriented boxes"),
                RuntimeWarning, stacklevel=3)
        self.model = native.Model(
            self.flat, goal, fing, kind=0 if hands else 1, iterations=iterations, tolerance=tolerance,
            n_steps_lookahead=L, wrong_press_termination=bool(kw["wrong_press_termination"]),
            disable_fingering_reward=self.disable_fingering_reward,
            disable_forearm_reward=bool(kw["disable_forearm_reward"]),
            energy_penalty_coef=float(kw["energy_penalty_coef"]), reward_type=reward_type,
            ot_fingering_reward=cfg["ot_fingering_reward"])
        self.num_envs, self.device = int(num_envs), torch.device("cuda", device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.batch = native.Batch(self.model, self.num_envs, device=device, stream=stream)
        import os
        self.randomize_hand_positions = bool(hands and kw["randomize_hand_positions"])
        if self.randomize_hand_positions:  # piano_with_shadow_hands.py:491-499, one generator stream per environment
            self.batch.randomize_hand_positions(2 if recompile_physics else 1, self.random_state.randint(0, 2**31 - 1))
        if os.environ.get("ROBOPLAY_BLOCK_ENVS"):   # launch shape under test (results do not depend on it)
            self.batch.set_block_envs(int(os.environ["ROBOPLAY_BLOCK_ENVS"]))
        N, dev = self.num_envs, self.device
        self.obs = torch.empty((N, self.model.obs_dim), dtype=torch.float64, device=dev)
        self.reward = torch.empty(N, dtype=torch.float64, device=dev)
        self.discount = torch.empty(N, dtype=torch.float64, device=dev)
        self.step_type = torch.empty(N, dtype=torch.uint8, device=dev)
        self.reward_terms = torch.empty((N, 5), dtype=torch.float64, device=dev)
        self._hands = hands
        self.obs_slices = self._obs_layout()

    # ------------------------------------------------------------------ specs
    def _obs_layout(self):
        hand_dofs = int(self.flat.arrays["hand_obs_dofid"].shape[1]) if self._hands else 0
        out = obs_layout(self._hands, self.n_steps_lookahead, self.disable_fingering_reward, hand_dofs)
        off = next(reversed(out.values())).stop
        assert off == self.model.obs_dim, (off, self.model.obs_dim)
        return out

    def action_spec(self) -> BoundedArray:
        cr = self.flat.actuator_ctrlrange
        lo, hi = np.append(cr[:, 0], 0.0), np.append(cr[:, 1], 1.0)
        names = "\t".join(list(self.flat.meta["actuator_names"]) + ["sustain"])
        return BoundedArray((self.model.act_dim,), np.float64, lo, hi, name=names)

    def observation_spec(self):
        return collections.OrderedDict((k, Array((s.stop - s.start,), np.float64, k)) for k, s in self.obs_slices.items())

    # ------------------------------------------------------------------ stepping
    def reset(self, mask=None):
        """FIRST observation for all envs (mask: uint8 CUDA tensor selecting the envs to reset)."""
        if mask is None and self._augmentations:
            self._change_midi()
        self.batch.reset(None if mask is None else mask.data_ptr(), self.obs.data_ptr())
        if mask is None:
            self.reward.zero_(); self.discount.fill_(1.0); self.step_type.fill_(int(StepType.FIRST))
        else:  # the reset envs show FIRST / 0 / 1 like a full reset; the others keep what their last step returned
            m = mask.to(self.torch.bool)
            self.reward.masked_fill_(m, 0.0); self.discount.masked_fill_(m, 1.0)
            self.step_type.masked_fill_(m, int(StepType.FIRST))
        return self.obs, self.reward, self.discount, self.step_type

    def _change_midi(self):
        """PianoWithShadowHands._maybe_change_midi + _reset_trajectory (piano_with_shadow_hands.py:151-165)."""
        from . import variations
        midi = variations.apply(self._augmentations, self._initial_midi, self.random_state)
        traj = trajectory.NoteTrajectory.from_seq(midi, self.control_timestep)
        if self._initial_buffer_time:
            traj = traj.add_initial_buffer_time(self._initial_buffer_time)
        self.midi, self.trajectory = midi, traj
        self._notes, self._sustains = traj.notes, traj.sustains
        self.goal_table, self.fingering_table = traj.goal_table(), traj.fingering_table()
        self.batch.set_song(self.goal_table, self.fingering_table)

    def step(self, actions):
        """actions: float64 CUDA tensor [N, act_dim] (contiguous). One launch; asynchronous."""
        t = self.torch
        if not (isinstance(actions, t.Tensor) and actions.is_cuda and actions.dtype == t.float64
                and actions.is_contiguous() and tuple(actions.shape) == (self.num_envs, self.model.act_dim)):
            raise ValueError(f"actions must be a contiguous float64 CUDA tensor of shape {(self.num_envs, self.model.act_dim)}")
        self.batch.step(actions.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(), self.discount.data_ptr(),
                        self.step_type.data_ptr(), self.reward_terms.data_ptr())
        return self.obs, self.reward, self.discount, self.step_type

    def step_host(self, actions: np.ndarray, copy: bool = True):
        """Host buffers in and out (pinned staging + copies inside the library): the e2e call.  Returns (obs, reward,
        discount, step_type, reward_terms) as numpy arrays; copy=False hands out views of the pinned output buffer
        (overwritten by the next call)."""
        return self.batch.step_host(actions, copy=copy)

    def get_musical_metrics(self):
        """Mean precision / recall / F1 (keys and sustain) over the envs that completed an episode:
        the batched form of MidiEvaluationWrapper.get_musical_metrics (wrappers/evaluation.py:95-110,
        deque_size=1), accumulated inside the step kernel."""
        m = self.batch.get_metrics()
        done = m["episodes"] > 0
        if not done.any():
            raise ValueError("No episode metrics available yet.")
        return {k: float(m[k][done].mean()) for k in
                ("precision", "recall", "f1", "sustain_precision", "sustain_recall", "sustain_f1")}

    def enable_midi(self, capacity: int = 4096):
        """Record the MIDI message stream of every env on the device (Piano.midi_module of the reference:
        models/piano/midi_module.py:47-98, one activation XOR per physics substep)."""
        self.batch.midi_enable(capacity)

    def get_all_midi_messages(self):
        """Per env, the messages of the running episode as (time, kind, midi_note) in the reference's order
        (MidiModule.get_all_midi_messages); read them before stepping past the end of the episode."""
        return self.batch.get_midi(self.flat.meta["timestep"])

    def uncovered_collider_flags(self):
        """bool [N]: during the last step a hand collider the narrowphase does not cover (meta['ignored_colliders'])
        came within reach of the piano — its contact, if any, was not generated."""
        return (self.batch.get_diag()["flags"] & 8) != 0

    def state_dict(self):
        """Checkpoint of the running batch: the device state (rp_checkpoint_save) plus the host-side song of the running
        episode; `load_state_dict` on an environment built with the same arguments resumes bit for bit."""
        return {"device": self.batch.checkpoint(), "goal": self.goal_table.copy(), "fingering": self.fingering_table.copy(),
                "random_state": self.random_state.get_state()}

    def load_state_dict(self, sd):
        if sd["goal"].shape != self.goal_table.shape or not (np.array_equal(sd["goal"], self.goal_table)
                                                              and np.array_equal(sd["fingering"], self.fingering_table)):
            self.goal_table, self.fingering_table = sd["goal"].copy(), sd["fingering"].copy()
            self.batch.set_song(self.goal_table, self.fingering_table)
        self.batch.restore(sd["device"])
        self.random_state.set_state(sd["random_state"])

    def get_state(self):
        return self.batch.get_state()

    def set_state(self, *a, **k):
        return self.batch.set_state(*a, **k)

    def close(self):
        self.batch.close()


class _Piano:
    """`env.task.piano` attributes read by the reference's wrappers (wrappers/evaluation.py:70-73)."""

    def __init__(self, env):
        self._env = env

    @property
    def activation(self):
        return self._env._batched.batch.get_aux()["activation"][0].astype(bool)

    @property
    def sustain_activation(self):
        return np.array([self._env._sustain >= 0.5])

    @property
    def midi_module(self):
        return self

    def get_all_midi_messages(self):
        """`env.task.piano.midi_module.get_all_midi_messages()` (needs `env.enable_midi()` before reset)."""
        return self._env._batched.get_all_midi_messages()[0]


class _RewardFn:
    def __init__(self):
        self.reward_terms = {}


class _Task:
    """`env.task` facade: `_notes`, `_sustains`, `piano`, `reward_fn.reward_terms`, `control_timestep`."""

    def __init__(self, env):
        self._b = env._batched
        self.piano = _Piano(env)
        self.reward_fn = _RewardFn()
        self.control_timestep = self._b.control_timestep

    # the song may change at every reset (augmentations): always the running episode's
    @property
    def _notes(self):
        return self._b._notes

    @property
    def _sustains(self):
        return self._b._sustains

    @property
    def midi(self):
        return self._b.midi


_TERM_NAMES = ("key_press_reward", "sustain_reward", "energy_reward", "fingering_reward", "forearm_reward")


class PianoEnv:
    """One environment with the ``dm_env.Environment`` call surface (reset / step / specs)."""

    def __init__(self, batched: BatchedPianoEnv):
        assert batched.num_envs == 1
        self._batched = batched
        self._sustain = 0.0
        self.task = _Task(self)
        self._reset_next = True

    def action_spec(self):
        return self._batched.action_spec()

    def observation_spec(self):
        return self._batched.observation_spec()

    def _obs_dict(self, obs_row):
        return collections.OrderedDict((k, obs_row[s].copy()) for k, s in self._batched.obs_slices.items())

    def reset(self) -> TimeStep:
        b = self._batched
        b.reset()
        b.torch.cuda.synchronize(b.device)
        self._reset_next = False
        self._sustain = 0.0
        return TimeStep(StepType.FIRST, None, None, self._obs_dict(b.obs[0].cpu().numpy()))

    def step(self, action) -> TimeStep:
        if self._reset_next:
            return self.reset()
        b = self._batched
        action = self.action_spec().validate(np.asarray(action, np.float64))
        obs, rew, dis, st, terms = b.step_host(action[None])
        self._sustain = float(action[-1])
        names = _TERM_NAMES if b._hands else ("key_press_reward",)
        enabled = [n for n in names if not (n == "forearm_reward" and b.disable_forearm_reward)]
        # the fourth slot carries the fingering reward or, when that one is disabled, the optimal-transport one
        rename = {"fingering_reward": "ot_fingering_reward"} if b.disable_fingering_reward else {}
        self.task.reward_fn.reward_terms = {rename.get(n, n): float(terms[0][names.index(n)]) for n in enabled}
        step_type = StepType(int(st[0]))
        if step_type == StepType.LAST:
            self._reset_next = True
        return TimeStep(step_type, float(rew[0]), float(dis[0]), self._obs_dict(obs[0]))

    def get_musical_metrics(self):
        return self._batched.get_musical_metrics()

    def enable_midi(self, capacity: int = 4096):
        self._batched.enable_midi(capacity)

    def close(self):
        self._batched.close()

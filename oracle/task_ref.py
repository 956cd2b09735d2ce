"""CPU restatement of the per-step task logic around the oracle physics — TEST INFRASTRUCTURE ONLY.

Follows, line by line, the reference's Python:
  composer.Environment.step order: before_step -> n_sub x (physics.step,
  Piano.after_substep) -> after_step -> observables -> reward -> termination;
  robopianist/suite/tasks/piano_with_shadow_hands.py:176-220 (before/after_step, termination),
  :251-369 (rewards, incl. the optimal-transport fingering reward), :371-412 (goal / fingering state), robopianist/models/piano/piano.py:178-192
  (key state), robopianist/models/piano/midi_module.py:40-98 (MIDI messages per substep, record_midi=True),
  robopianist/suite/tasks/self_actuated_piano.py:38-57,143-243,
  dm_control.utils.rewards.tolerance (gaussian sigmoid, value_at_margin=0.1).
PARITY UNPINNED for the physics underneath (see oracle.c).
"""
from __future__ import annotations

import numpy as np

from . import oracle as O

KEY_THRESHOLD = 0.00872665
SUSTAIN_THRESHOLD = 0.5
FIRST, MID, LAST = 0, 1, 2


def tolerance(x, bounds, margin):
    x = np.asarray(x, dtype=np.float64)
    lo, hi = bounds
    in_bounds = np.logical_and(lo <= x, x <= hi)
    d = np.where(x < lo, lo - x, x - hi) / margin
    scale = np.sqrt(-2 * np.log(0.1))
    return np.where(in_bounds, 1.0, np.exp(-0.5 * (d * scale) ** 2))


def prf_binary(y_true, y_pred):
    """precision_recall_fscore_support(average="binary", zero_division=1) as called by the reference's
    MidiEvaluationWrapper (wrappers/evaluation.py:140-147): restated so the tests can pin it to sklearn."""
    y_true, y_pred = np.asarray(y_true) > 0.5, np.asarray(y_pred) > 0.5
    tp = float((y_true & y_pred).sum()); fp = float((~y_true & y_pred).sum()); fn = float((y_true & ~y_pred).sum())
    div = lambda a, b: a / b if b > 0 else 1.0
    return div(tp, tp + fp), div(tp, tp + fn), div(2 * tp, 2 * tp + fp + fn)


class RefEnv:
    """One environment: oracle physics + the reference's task logic (dm_env semantics)."""

    def __init__(self, fm, goal, fingering, kind=0, n_steps_lookahead=1, wrong_press_termination=False,
                 disable_fingering_reward=False, disable_forearm_reward=False, energy_penalty_coef=5e-3,
                 reward_type=0, maxcon=24, iterations=None, tolerance_=None, ot_fingering_reward=False):
        self.fm, self.kind = fm, kind
        self.goal = np.asarray(goal, np.float64)
        self.fingering = np.asarray(fingering)
        self.T, self.L = goal.shape[0], n_steps_lookahead
        self.wrong_press_termination = wrong_press_termination
        self.disable_fingering_reward = disable_fingering_reward
        self.disable_forearm_reward = disable_forearm_reward
        self.energy_coef, self.reward_type = energy_penalty_coef, reward_type
        self.ot_fingering_reward = ot_fingering_reward and disable_fingering_reward and kind == 0
        self.om = O.OracleModel(fm, iterations, tolerance_)
        self.om.set_option(maxcon=maxcon)
        self.d = O.OracleData(self.om)
        self.qmax = fm.key_qmax
        self.nu = fm.meta["nu"]
        self._need_reset = True
        self.reward_terms = {}
        self.record_midi = False
        self.midi_messages = []

    # ---------------------------------------------------------------- helpers
    def _forearms_touch(self):
        fm, d = self.fm, self.d
        fore = [[g for g in range(fm.meta["ngeom"]) if fm.geom_forearm[g] and fm.geom_hand[g] == h] for h in (0, 1)]
        if not fore[0] or not ((fm.geom_contype[fore[0][0]] & fm.geom_conaffinity[fore[1][0]])
                               or (fm.geom_contype[fore[1][0]] & fm.geom_conaffinity[fore[0][0]])):
            return False

        def box(g):  # centre, axes (columns), half sizes; cylinders never occur on the forearm
            return d.geom_xpos[g], d.geom_xmat[g].reshape(3, 3), fm.geom_size[g]

        def overlap(a, b):  # separating-axis theorem, plain form: project both boxes on the 15 candidate axes
            (ca, Ra, ha), (cb, Rb, hb) = a, b
            axes = [Ra[:, i] for i in range(3)] + [Rb[:, j] for j in range(3)]
            axes += [np.cross(Ra[:, i], Rb[:, j]) for i in range(3) for j in range(3)]
            for ax in axes:
                n = np.linalg.norm(ax)
                if n < 1e-9:
                    continue
                ax = ax / n
                ra = sum(ha[i] * abs(ax @ Ra[:, i]) for i in range(3))
                rb = sum(hb[j] * abs(ax @ Rb[:, j]) for j in range(3))
                if abs(ax @ (cb - ca)) > ra + rb:
                    return False
            return True

        return any(overlap(box(g0), box(g1)) for g0 in fore[0] for g1 in fore[1])

    def _update_key_state(self):
        d = self.d
        if self.kind == 1:
            mid = 0.5 * (self.fm.actuator_ctrlrange[:, 0] + self.fm.actuator_ctrlrange[:, 1])
            self.activation = d.ctrl >= mid
        else:
            self.state = np.clip(d.qpos[:88], 0.0, self.qmax)
            self.normalized = self.state / self.qmax
            self.activation = np.abs(self.state - self.qmax) <= KEY_THRESHOLD
        self.sustain_activation = self.sustain_state >= SUSTAIN_THRESHOLD

    def _midi_after_substep(self):
        """MidiModule.after_substep (midi_module.py:47-98): note-ons (ascending key), note-offs, pedal."""
        self._substeps += 1
        t = self._substeps * self.fm.meta["timestep"]
        act, prev = np.asarray(self.activation, bool), self._prev_activation
        change = act ^ prev
        for k in np.flatnonzero(change & ~prev):
            self.midi_messages.append((t, "note_on", int(k) + 21))
        for k in np.flatnonzero(change & ~act):
            self.midi_messages.append((t, "note_off", int(k) + 21))
        sus, psus = bool(self.sustain_activation), self._prev_sustain_activation
        if (sus != psus) and not psus:
            self.midi_messages.append((t, "sustain_on", None))
        if (sus != psus) and not sus:
            self.midi_messages.append((t, "sustain_off", None))
        self._prev_activation, self._prev_sustain_activation = act.copy(), sus

    def _update_goal_state(self):
        if self.t_idx == self.T:
            return
        g = np.zeros((self.L + 1, 89))
        for i, t in enumerate(range(self.t_idx, min(self.t_idx + self.L + 1, self.T))):
            g[i] = self.goal[t]
        self.goal_state = g
        f = self.fingering[self.t_idx]
        keys = np.flatnonzero(self.goal[self.t_idx, :88])
        self.rh_keys = [(k, int(f[k])) for k in keys if 0 <= f[k] < 5]
        self.lh_keys = [(k, int(f[k]) - 5) for k in keys if f[k] >= 5]
        fs = np.zeros((2, 5))
        for k, fi in self.rh_keys:
            fs[0, fi] = 1
        for k, fi in self.lh_keys:
            fs[1, fi] = 1
        self.fingering_state = fs

    def _observation(self):
        self._update_goal_state()
        d, fm = self.d, self.fm
        if self.kind == 0:
            parts = [self.normalized, [self.sustain_state], d.qpos[fm.hand_obs_dofid[0]], d.qpos[fm.hand_obs_dofid[1]],
                     self.goal_state.ravel()]
            if not self.disable_fingering_reward:
                parts.append(self.fingering_state.ravel())
        else:
            parts = [self.activation.astype(np.float64), [float(self.sustain_activation)], self.goal_state.ravel()]
        return np.concatenate([np.asarray(p, np.float64).ravel() for p in parts])

    # ---------------------------------------------------------------- dm_env API
    def reset(self):
        self.d.reset()
        self.t_idx = 0
        self.sustain_state = 0.0
        self._update_key_state()
        self._need_reset = False
        self.last_actuator_force = np.zeros(self.nu)
        self._key_prf, self._sustain_prf = [], []  # MidiEvaluationWrapper.reset (wrappers/evaluation.py:90-93)
        # MidiModule.initialize_episode (midi_module.py:40-45)
        self._prev_activation = np.zeros(88, dtype=bool)
        self._prev_sustain_activation = False
        self.midi_messages = []
        self._substeps = 0
        return FIRST, None, None, self._observation()

    def step(self, action):
        if self._need_reset:
            return self.reset()
        d, fm = self.d, self.fm
        action = np.asarray(action, np.float64)
        self.sustain_state = float(action[-1])
        if not self.record_midi:
            d.control_step(action[:-1])  # nsub x (mj_step2; mj_step1): legacy_step=True
        else:
            for _ in range(self.om.nsubstep):
                d.control_step(action[:-1], 1)
                self._update_key_state()  # Piano.after_substep (piano.py:154-162)
                self._midi_after_substep()
        self._update_key_state()
        # after_step
        self.t_idx += 1
        should_terminate = (self.t_idx - 1) == self.T - 1
        goal_current = self.goal_state[0].copy()
        rh_cur, lh_cur = list(self.rh_keys), list(self.lh_keys)
        failure = bool(self.activation[np.flatnonzero(1 - goal_current[:-1])].any()) if self.kind == 0 else False
        obs = self._observation()
        # reward
        if self.kind == 0:
            on = np.flatnonzero(goal_current[:-1])
            r_key = 0.0
            if on.size > 0:
                actual = self.state / self.qmax
                r_key += 0.5 * tolerance(goal_current[:-1][on] - actual[on], (0, 0.05), 0.5).mean()
            off = np.flatnonzero(1 - goal_current[:-1])
            r_key += 0.5 * (1 - float(self.activation[off].any()))
            r_sus = float(tolerance(goal_current[-1] - float(self.sustain_activation), (0, 0.05), 0.5))
            # energy: actuatorfrc from the last mj_step2 (pre-integration), actuatorvel from the last mj_step1
            power = np.abs(d.actuator_force) * np.abs(d.actuator_velocity)
            r_en = -self.energy_coef * power[:22].sum() - self.energy_coef * power[22:].sum()
            terms = dict(key_press_reward=r_key, sustain_reward=r_sus, energy_reward=r_en)
            if not self.disable_fingering_reward:
                dists = []
                for hand, keys in ((0, rh_cur), (1, lh_cur)):
                    for k, fi in keys:
                        g = fm.key_geomid[k]
                        kp = d.geom_xpos[g].copy()
                        kp[2] += 0.5 * fm.geom_size[g, 2]
                        kp[0] += 0.35 * fm.geom_size[g, 0]
                        dists.append(np.linalg.norm(kp - d.site_xpos[hand * 5 + fi]))
                terms["fingering_reward"] = float(tolerance(np.array(dists), (0, 0.01), 0.1).mean()) if dists else 0.0
            elif self.ot_fingering_reward:
                # _compute_ot_fingering_reward (piano_with_shadow_hands.py:333-369): left-hand fingertips first
                from scipy.optimize import linear_sum_assignment
                keys_to_press = np.flatnonzero(goal_current[:-1])
                if keys_to_press.shape[0] == 0:
                    terms["ot_fingering_reward"] = 1.0
                else:
                    tips = [d.site_xpos[5 + i].copy() for i in range(5)] + [d.site_xpos[i].copy() for i in range(5)]
                    key_pos = []
                    for k in keys_to_press:
                        g = fm.key_geomid[k]
                        kp = d.geom_xpos[g].copy()
                        kp[-1] += 0.5 * fm.geom_size[g, 2]
                        kp[0] += 0.35 * fm.geom_size[g, 0]
                        key_pos.append(kp)
                    dist = np.full((len(tips), len(key_pos)), 100.0)
                    for i, f in enumerate(tips):
                        for j, kpos in enumerate(key_pos):
                            dist[i, j] = np.linalg.norm(kpos - f)
                    ri, ci = linear_sum_assignment(dist)
                    terms["ot_fingering_reward"] = float(np.mean(tolerance(dist[ri, ci], (0, 0.01), 0.1)))
            if not self.disable_forearm_reward:
                # _compute_forearm_reward (piano_with_shadow_hands.py:251-259): 0 when a geom of one root body touches a
                # geom of the other.  With disable_hand_collisions=True the hands cannot touch each other.  Otherwise the
                # forearm colliders (a box and a mesh, outside the narrowphase's coverage) are tested as oriented boxes
                # (box / mesh bounding box), an approximation of MuJoCo's convex contact test.
                terms["forearm_reward"] = 0.0 if self._forearms_touch() else 0.5
            reward = sum(terms.values())
        else:
            pred = np.concatenate([self.activation.astype(np.float64), [float(self.sustain_activation)]])
            if self.reward_type == 0:
                reward = -np.sqrt(np.sum((pred - goal_current) ** 2))
            else:
                reward = np.sum(goal_current * np.log(pred + 1e-6) + (1 - goal_current) * np.log(1 - pred + 1e-6))
            terms = dict(key_press_reward=reward)
        self.reward_terms = terms
        discount = 1.0
        step_type = MID
        if should_terminate:
            step_type = LAST
        elif self.kind == 0 and self.wrong_press_termination and failure:
            discount, step_type = 0.0, LAST
        # MidiEvaluationWrapper.step (wrappers/evaluation.py:67-88): activation after this step vs. the
        # ground truth of this step index (`_notes[i]` == the goal emitted with the previous observation)
        self._key_prf.append(prf_binary(goal_current[:-1], self.activation))
        self._sustain_prf.append(prf_binary([goal_current[-1]], [float(self.sustain_activation)]))
        if step_type == LAST:
            self._need_reset = True
            k, su = np.mean(self._key_prf, axis=0), np.mean(self._sustain_prf, axis=0)
            self.musical_metrics = dict(precision=k[0], recall=k[1], f1=k[2], sustain_precision=su[0],
                                        sustain_recall=su[1], sustain_f1=su[2])
        return step_type, float(reward), discount, obs

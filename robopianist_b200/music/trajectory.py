"""Song -> per-control-step goal tables.

Restates ``NoteTrajectory.seq_to_trajectory`` (robopianist/music/midi_file.py:315-362) on top of
the piano-roll rules of ``sequence_to_pianoroll`` (robopianist/music/piano_roll.py:59-204, called
with ``onset_window=0``): the roll has ``int(total_time*fps + 1)`` frames; a note covers frames
``[int(start*fps), max(start_frame+1, ceil(end*fps)))``; a note that re-strikes a key that was
already active in the previous frame is dropped for its onset frame (so the key is released for
one step); sustain follows CC64 events stored as ``value+1``.

The dense outputs (``goal[T, 89]`` uint8 and ``fingering`` tables) are what the device kernels
stream; they are the batched form of ``_update_goal_state`` / ``_update_fingering_state``
(robopianist/suite/tasks/piano_with_shadow_hands.py:371-412).
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Tuple

import numpy as np

from .sequence import MAX_MIDI_PITCH_PIANO, MIN_MIDI_PITCH_PIANO, NoteSeq

SUSTAIN_CC = 64
MAX_CC_VALUE = 127
MAX_VELOCITY = 127
NUM_KEYS = 88


@dataclasses.dataclass(frozen=True)
class PianoNote:
    number: int
    velocity: int
    key: int
    fingering: int = -1


def _frames(start_time: float, end_time: float, fps: float) -> Tuple[int, int]:
    start_frame = int(start_time * fps)
    end_frame = int(math.ceil(end_time * fps))
    return start_frame, max(start_frame + 1, end_frame)


@dataclasses.dataclass
class NoteTrajectory:
    dt: float
    notes: List[List[PianoNote]]
    sustains: List[int]

    def __post_init__(self):
        if self.dt <= 0:
            raise ValueError("dt must be positive.")
        if len(self.notes) != len(self.sustains):
            raise ValueError("notes and sustains must have the same length.")

    def __len__(self):
        return len(self.notes)

    @classmethod
    def from_seq(cls, seq: NoteSeq, dt: float) -> "NoteTrajectory":
        fps = 1 / dt
        n_frames = int(seq.total_time * fps + 1)
        active_vel = np.zeros((n_frames, 128), np.float32)
        onsets = np.zeros((n_frames, 128), np.float32)
        fingerings = np.full((n_frames, 128), -1.0, np.float32)
        for note in sorted(seq.notes, key=lambda n: n.start_time):
            s, e = _frames(note.start_time, note.end_time, fps)
            onset_end = min(n_frames, s + 1)  # onset window 0
            onsets[s:onset_end, note.pitch] = 1.0
            active_vel[s:e, note.pitch] = np.float32(note.velocity / MAX_VELOCITY)
            fingerings[s:e, note.pitch] = note.part
        onset_vel = active_vel * onsets
        cc = np.zeros((n_frames, 128), np.int32)
        for c in seq.control_changes:
            frame = int(c.time * fps)
            if frame < n_frames:
                cc[frame, c.control_number] = c.control_value + 1

        notes: List[List[PianoNote]] = []
        for t in range(n_frames):
            row = []
            for pitch in np.nonzero(active_vel[t])[0]:
                if t > 0 and active_vel[t - 1][pitch] and onset_vel[t][pitch]:
                    continue  # re-struck key: release for one frame
                if not MIN_MIDI_PITCH_PIANO <= pitch <= MAX_MIDI_PITCH_PIANO:
                    raise ValueError(f"MIDI pitch {pitch} is outside the piano range")
                vel = int(round(float(active_vel[t][pitch]) * MAX_VELOCITY))
                row.append(PianoNote(int(pitch), vel, int(pitch) - MIN_MIDI_PITCH_PIANO,
                                     int(fingerings[t, pitch])))
            notes.append(row)
        sustains, prev = [], 0
        for t in range(n_frames):
            ev = cc[t, SUSTAIN_CC]
            if 1 <= ev <= SUSTAIN_CC:
                s = 0
            elif SUSTAIN_CC + 1 <= ev <= MAX_CC_VALUE + 1:
                s = 1
            else:
                s = prev
            sustains.append(s)
            prev = s
        return cls(dt=dt, notes=notes, sustains=sustains)

    def add_initial_buffer_time(self, initial_buffer_time: float) -> "NoteTrajectory":
        if initial_buffer_time < 0.0:
            raise ValueError("initial_buffer_time must be non-negative.")
        for _ in range(int(round(initial_buffer_time / self.dt))):
            self.notes.insert(0, [])
            self.sustains.insert(0, 0)
        return self

    # ---- dense device tables ----------------------------------------------------------------
    def goal_table(self) -> np.ndarray:
        """uint8 [T, 89]: columns 0..87 key goals, column 88 sustain."""
        T = len(self.notes)
        g = np.zeros((T, NUM_KEYS + 1), np.uint8)
        for t, row in enumerate(self.notes):
            for n in row:
                g[t, n.key] = 1
            g[t, NUM_KEYS] = self.sustains[t]
        return g

    def fingering_table(self) -> np.ndarray:
        """int8 [T, 88]: finger 0-4 (right) / 5-9 (left) assigned to a goal key, -1 otherwise."""
        T = len(self.notes)
        f = np.full((T, NUM_KEYS), -1, np.int8)
        for t, row in enumerate(self.notes):
            for n in row:
                f[t, n.key] = n.fingering
        return f

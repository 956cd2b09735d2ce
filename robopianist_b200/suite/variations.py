"""Song augmentations (robopianist/suite/variations.py:27-184) over ``NoteSeq``, without dm_control.

Same call signature as ``dm_control.composer.variation.Variation`` (``initial_value``, ``current_value``,
``random_state``) and the same sequence of draws from the ``numpy.random.RandomState``, so that a seeded reference run
and a seeded run here pick the same songs: ``distributions.UniformChoice`` draws ``random_state.choice(choices)``,
``distributions.Uniform(lo, hi)`` draws ``random_state.uniform(lo, hi)``.  Applied by the environment at every episode
start, in order, starting from the initial song (piano_with_shadow_hands.py:461-470).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from ..music import library
from ..music.sequence import MAX_MIDI_PITCH_PIANO, MIN_MIDI_PITCH_PIANO, NoteSeq


def _need_initial(initial_value):
    if initial_value is None or not isinstance(initial_value, NoteSeq):
        raise ValueError("Expected `initial_value` to be provided and be a NoteSeq.")
    return initial_value


class MidiSelect:
    """Randomly select a song from the registry (variations.py:27-46)."""

    def __init__(self, midi_names: Sequence[str] = ()) -> None:
        self._midi_names = list(midi_names)

    def __call__(self, initial_value=None, current_value=None, random_state=None) -> NoteSeq:
        del initial_value, current_value
        random = random_state or np.random
        return library.load(str(random.choice(self._midi_names)))


class MidiTemporalStretch:
    """With probability `prob`, stretch by a factor drawn from U[1 - r, 1 + r] (variations.py:49-82)."""

    def __init__(self, prob: float, stretch_range: float) -> None:
        self._prob, self._range = prob, stretch_range

    def __call__(self, initial_value=None, current_value=None, random_state=None) -> NoteSeq:
        del current_value
        random = random_state or np.random
        if random.uniform(0.0, 1.0) > self._prob:
            return _need_initial(initial_value)
        return initial_value.stretch(1.0 + random.uniform(-self._range, self._range))


class _Shift:
    def __init__(self, prob: float, rng: int, unit: int, what: str) -> None:
        if not isinstance(rng, int):
            raise ValueError(f"`{what}` must be an integer.")
        self._prob, self._rng, self._unit = prob, rng, unit

    def __call__(self, initial_value=None, current_value=None, random_state=None) -> NoteSeq:
        del current_value
        random = random_state or np.random
        if random.uniform(0.0, 1.0) > self._prob:
            return _need_initial(initial_value)
        if self._rng == 0:
            return initial_value
        pitches = [n.pitch for n in initial_value.notes]
        low = max(MIN_MIDI_PITCH_PIANO - min(pitches), -self._rng * self._unit)
        high = min(MAX_MIDI_PITCH_PIANO - max(pitches), self._rng * self._unit)
        shift = random.randint(low // self._unit, high // self._unit + 1)
        return initial_value if shift == 0 else initial_value.transpose(shift * self._unit)


class MidiPitchShift(_Shift):
    """With probability `prob`, transpose by up to `shift_range` semitones within the piano's range (variations.py:85-133)."""

    def __init__(self, prob: float, shift_range: int) -> None:
        super().__init__(prob, shift_range, 1, "shift_range")


class MidiOctaveShift(_Shift):
    """With probability `prob`, transpose by up to `octave_range` octaves within the piano's range (variations.py:136-184)."""

    def __init__(self, prob: float, octave_range: int) -> None:
        super().__init__(prob, octave_range, 12, "octave_range")


def apply(augmentations, initial_midi: NoteSeq, random_state) -> NoteSeq:
    """piano_with_shadow_hands.py:461-470 (_maybe_change_midi): every variation sees the result of the previous one."""
    midi = initial_midi
    for var in augmentations or ():
        midi = var(initial_value=midi, random_state=random_state)
    return midi

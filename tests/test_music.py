"""Goal-table generation vs the reference's structural pins:
midi_file_test.py:175-211 (re-strike gap, sustain decoding), piano_with_shadow_hands_test.py:152-226
(goal / fingering lookups), library.py:346-385 (Twinkle fingering order), §8(a) episode lengths."""
import numpy as np
import pytest

from robopianist_b200.music import library, trajectory
from robopianist_b200.music.sequence import ControlChange, NoteSeq


def test_episode_lengths():
    for name, T in (("CMajorScaleTwoHands", 151), ("TwinkleTwinkleLittleStar", 161), ("CMajorChordProgressionTwoHands", 81)):
        assert len(trajectory.NoteTrajectory.from_seq(library.load(name), 0.05)) == T
    tw = library.load("TwinkleTwinkleRousseau")
    assert len(trajectory.NoteTrajectory.from_seq(tw, 0.05)) == 451
    assert len(trajectory.NoteTrajectory.from_seq(tw.trim_silence(), 0.05)) == 158  # == len(recorded action replay)


def test_cmajor_two_goal_keys_every_step():
    traj = trajectory.NoteTrajectory.from_seq(library.load("CMajorScaleTwoHands"), 0.05)
    g = traj.goal_table()
    counts = g[:, :88].sum(axis=1)
    # the last frame (t = total_time exactly) is empty; re-struck notes never occur in the scale
    assert (counts[:150] == 2).all() and counts[150] == 0
    f = traj.fingering_table()
    assert ((f >= 0) == (g[:, :88] == 1)).all()


def test_twinkle_rousseau_fingering_and_has_fingering():
    seq = library.load("TwinkleTwinkleRousseau")
    assert seq.has_fingering() and seq.n_notes == 34
    order = sorted(range(34), key=lambda i: seq.notes[i].start_time)
    assert [seq.notes[i].part for i in order][:6] == [1, 9, 5, 0, 3, 6]
    assert [seq.notes[i].pitch for i in order][:4] == [72, 48, 60, 72]  # C5 C3 C4 C5


def _two_note_seq(gap):
    seq = NoteSeq()
    seq.add(60, 0.0, 0.5, 80, 1)
    seq.add(60, 0.5 + gap, 1.0, 80, 1)
    seq.total_time = 1.0
    return seq


def test_restrike_releases_key_for_one_step():
    # consecutive same-pitch notes: the onset frame of the second note is dropped (midi_file.py:334-342)
    traj = trajectory.NoteTrajectory.from_seq(_two_note_seq(0.0), 0.05)
    on = [bool(row) for row in traj.notes]
    assert on[:10] == [True] * 10 and on[10] is False and on[11] is True


def test_sustain_decoding():
    seq = _two_note_seq(0.0)
    seq.control_changes = [ControlChange(0.2, 64, 127), ControlChange(0.6, 64, 0)]
    traj = trajectory.NoteTrajectory.from_seq(seq, 0.05)
    s = traj.sustains
    assert s[:4] == [0, 0, 0, 0] and all(v == 1 for v in s[4:12]) and all(v == 0 for v in s[12:])


def test_initial_buffer_and_stretch_transpose():
    seq = library.load("CMajorScaleOneHand", stretch=2.0, shift=1)
    assert seq.total_time == pytest.approx(15.0) and seq.notes[0].pitch == 73
    traj = trajectory.NoteTrajectory.from_seq(library.load("CMajorScaleOneHand"), 0.05).add_initial_buffer_time(0.5)
    assert len(traj) == 161 and traj.notes[9] == [] and traj.notes[10]
    with pytest.raises(ValueError):
        library.load("CMajorScaleOneHand").stretch(0)
    with pytest.raises(KeyError):
        library.load("NoSuchSong")

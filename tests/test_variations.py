"""Song augmentations and the PIG .proto reader (suite/variations.py:27-184, music/midi_file.py:175-201 of the reference)."""
import numpy as np
import pytest

from robopianist_b200.music import library, proto
from robopianist_b200.music.sequence import NoteSeq
from robopianist_b200.suite import variations as V


def _seq():
    s = NoteSeq()
    s.add(60, 0.0, 0.5, 80, 1); s.add(64, 0.5, 1.0, 90, 6); s.add(67, 1.0, 1.75, 100, 2)
    s.total_time = 1.75
    return s


def test_temporal_stretch_follows_the_reference_draws():
    rs = np.random.RandomState(3)
    out = V.MidiTemporalStretch(prob=1.0, stretch_range=0.2)(initial_value=_seq(), random_state=rs)
    rs2 = np.random.RandomState(3)
    rs2.uniform(0.0, 1.0)                       # the coin
    f = 1.0 + rs2.uniform(-0.2, 0.2)            # distributions.Uniform(-r, r)
    assert [n.start_time for n in out.notes] == [n.start_time * f for n in _seq().notes]
    assert abs(out.total_time - 1.75 * f) < 1e-15
    same = V.MidiTemporalStretch(prob=0.0, stretch_range=0.2)(initial_value=_seq(), random_state=np.random.RandomState(0))
    assert [n.start_time for n in same.notes] == [0.0, 0.5, 1.0]
    with pytest.raises(ValueError):
        V.MidiTemporalStretch(prob=0.0, stretch_range=0.1)(initial_value=None, random_state=np.random.RandomState(0))


def test_pitch_and_octave_shift_follow_the_reference_formula():
    """variations.py:120-133 / 171-184: the shift is drawn with random.randint(low // unit, high // unit + 1) from bounds
    clipped to the piano's range (Python floor division: an octave shift can round the lower bound below the range, in
    which case transpose drops the notes that leave the keyboard, as note_seq's transpose_note_sequence does)."""
    for cls, kw, unit in ((V.MidiPitchShift, "shift_range", 1), (V.MidiOctaveShift, "octave_range", 12)):
        seen = set()
        for seed in range(200):
            out = cls(prob=1.0, **{kw: 100 if unit == 1 else 10})(initial_value=_seq(), random_state=np.random.RandomState(seed))
            rs = np.random.RandomState(seed)
            rs.uniform(0.0, 1.0)
            r = (100 if unit == 1 else 10) * unit
            low, high = max(21 - 60, -r), min(108 - 67, r)
            shift = rs.randint(low // unit, high // unit + 1) * unit
            want = [p + shift for p in (60, 64, 67) if 21 <= p + shift <= 108]
            assert [n.pitch for n in out.notes] == want
            seen.add(shift)
        assert len(seen) > (40 if unit == 1 else 5)
    with pytest.raises(ValueError):
        V.MidiPitchShift(prob=1.0, shift_range=1.5)
    assert V.MidiPitchShift(prob=1.0, shift_range=0)(initial_value=_seq(), random_state=np.random.RandomState(0)).notes[0].pitch == 60
    assert V.MidiOctaveShift(prob=0.0, octave_range=2)(initial_value=_seq(), random_state=np.random.RandomState(0)).notes[0].pitch == 60


def test_select_and_chained_application():
    names = ["CMajorScaleTwoHands", "TwinkleTwinkleLittleStar"] if "TwinkleTwinkleLittleStar" in library.MIDI_NAME_TO_CALLABLE else list(library.MIDI_NAME_TO_CALLABLE)[:2]
    rs = np.random.RandomState(5)
    picked = V.MidiSelect(names)(random_state=rs)
    assert picked.n_notes == library.load(str(np.random.RandomState(5).choice(names))).n_notes
    out = V.apply([V.MidiOctaveShift(1.0, 1), V.MidiTemporalStretch(1.0, 0.1)], _seq(), np.random.RandomState(1))
    assert out.n_notes == 3 and (out.notes[0].pitch - 60) % 12 == 0
    assert V.apply(None, _seq(), np.random.RandomState(1)).total_time == 1.75


def test_note_sequence_proto_round_trip(tmp_path):
    s = _seq()
    from robopianist_b200.music.sequence import ControlChange
    s.control_changes.append(ControlChange(0.25, 64, 127)); s.control_changes.append(ControlChange(1.5, 64, 0))
    data = proto.serialize_note_sequence(s)
    # wire format spot check: first record is field 8 (notes), length-delimited
    assert data[0] == (8 << 3 | 2)
    back = proto.parse_note_sequence(data)
    assert [(n.pitch, n.velocity, n.start_time, n.end_time, n.part) for n in back.notes] == \
           [(n.pitch, n.velocity, n.start_time, n.end_time, n.part) for n in s.notes]
    assert [(c.time, c.control_number, c.control_value) for c in back.control_changes] == [(0.25, 64, 127), (1.5, 64, 0)]
    assert back.total_time == 1.75 and back.has_fingering()
    p = tmp_path / "song.proto"
    p.write_bytes(data)
    assert library.load(p).n_notes == 3
    # unknown fields (ids, tempos, ...) are skipped
    extra = bytes([1 << 3 | 2, 3]) + b"abc" + bytes([4 << 3 | 0, 220, 1]) + data
    assert proto.parse_note_sequence(extra).n_notes == 3
    with pytest.raises(ValueError):
        library.load(tmp_path / "song.xyz")

"""A dependency-free stand-in for the slice of ``note_seq.NoteSequence`` the hot path needs.

The reference keeps songs as ``note_seq`` protobufs (robopianist/music/midi_file.py:166-292);
only notes (pitch, start, end, velocity, part=fingering), CC events and ``total_time`` feed the
goal tables, so that is all this class carries.  Also contains a small Standard-MIDI-File reader
that restates what ``note_seq.midi_io`` + ``pretty_midi`` do for a type-0/1 file: tempo map,
tick -> seconds, note-on/off pairing (notes are emitted in note-off order per instrument, as
pretty_midi does), CC events, ``total_time`` = latest note end.
"""
from __future__ import annotations

import copy
import dataclasses
import struct
from pathlib import Path
from typing import List

MIN_MIDI_PITCH_PIANO = 21
MAX_MIDI_PITCH_PIANO = 108


@dataclasses.dataclass
class Note:
    pitch: int
    start_time: float
    end_time: float
    velocity: int = 80
    part: int = 0  # fingering 0-4 right hand (thumb..pinky), 5-9 left hand, as in the reference


@dataclasses.dataclass
class ControlChange:
    time: float
    control_number: int
    control_value: int


@dataclasses.dataclass
class NoteSeq:
    notes: List[Note] = dataclasses.field(default_factory=list)
    control_changes: List[ControlChange] = dataclasses.field(default_factory=list)
    total_time: float = 0.0
    title: str = ""
    artist: str = ""

    def add(self, pitch, start_time, end_time, velocity=80, part=0):
        self.notes.append(Note(int(pitch), float(start_time), float(end_time), int(velocity), int(part)))

    # --- MidiFile methods of the reference (midi_file.py:203-263) -----------------------
    def stretch(self, factor: float) -> "NoteSeq":
        if factor <= 0:
            raise ValueError("factor must be positive.")
        out = copy.deepcopy(self)
        if factor == 1.0:
            return out
        for n in out.notes:
            n.start_time *= factor
            n.end_time *= factor
        for c in out.control_changes:
            c.time *= factor
        out.total_time *= factor
        return out

    def transpose(self, amount: int) -> "NoteSeq":
        out = copy.deepcopy(self)
        if amount == 0:
            return out
        kept = []
        for n in out.notes:
            n.pitch += amount
            if MIN_MIDI_PITCH_PIANO <= n.pitch <= MAX_MIDI_PITCH_PIANO:
                kept.append(n)
        out.notes = kept
        return out

    def trim_silence(self) -> "NoteSeq":
        """Subsequence [notes[0].start, notes[-1].end) shifted to start at 0 (midi_file.py:232-238)."""
        start = self.notes[0].start_time
        end = self.notes[-1].end_time
        out = copy.deepcopy(self)
        out.notes = []
        for n in self.notes:
            if n.start_time < start or n.start_time >= end:
                continue
            m = copy.copy(n)
            m.start_time = n.start_time - start
            m.end_time = min(n.end_time, end) - start
            out.notes.append(m)
        out.control_changes = [ControlChange(c.time - start, c.control_number, c.control_value)
                               for c in self.control_changes if start <= c.time < end]
        out.total_time = end - start
        return out

    def has_fingering(self) -> bool:
        parts = {n.part for n in self.notes}
        return len(parts) > 1 and any(p != 0 for p in parts)

    @property
    def duration(self) -> float:
        return self.total_time

    @property
    def n_notes(self) -> int:
        return len(self.notes)


# ---------------------------------------------------------------------------------------------
# Standard MIDI file reader
# ---------------------------------------------------------------------------------------------
def _read_vlq(data: bytes, i: int):
    v = 0
    while True:
        b = data[i]
        i += 1
        v = (v << 7) | (b & 0x7F)
        if not b & 0x80:
            return v, i


def _parse_tracks(data: bytes):
    if data[:4] != b"MThd":
        raise RuntimeError("not a MIDI file")
    hlen, fmt, ntrk, division = struct.unpack(">IHHH", data[4:14])
    if division & 0x8000:
        raise RuntimeError("SMPTE time division is not supported")
    i = 8 + hlen
    tracks = []
    for _ in range(ntrk):
        if data[i:i + 4] != b"MTrk":
            raise RuntimeError("bad track chunk")
        (tlen,) = struct.unpack(">I", data[i + 4:i + 8])
        j, end = i + 8, i + 8 + tlen
        tick, status, events = 0, 0, []
        while j < end:
            dt, j = _read_vlq(data, j)
            tick += dt
            b = data[j]
            if b == 0xFF:  # meta
                mtype = data[j + 1]
                ln, j = _read_vlq(data, j + 2)
                events.append((tick, "meta", mtype, data[j:j + ln]))
                j += ln
            elif b in (0xF0, 0xF7):
                ln, j = _read_vlq(data, j + 1)
                j += ln
            else:
                if b & 0x80:
                    status = b
                    j += 1
                hi, ch = status & 0xF0, status & 0x0F
                if hi in (0xC0, 0xD0):
                    events.append((tick, "ch", hi, ch, data[j], 0))
                    j += 1
                else:
                    events.append((tick, "ch", hi, ch, data[j], data[j + 1]))
                    j += 2
        tracks.append(events)
        i = end
    return division, tracks


def read_midi(path) -> NoteSeq:
    data = Path(path).read_bytes()
    division, tracks = _parse_tracks(data)
    # tempo map (all tracks), default 120 qpm
    tempo_events = sorted((ev[0], int.from_bytes(ev[3], "big")) for tr in tracks for ev in tr
                          if ev[1] == "meta" and ev[2] == 0x51)
    if not tempo_events or tempo_events[0][0] != 0:
        tempo_events.insert(0, (0, 500000))
    # cumulative seconds at each tempo change
    seg = []
    t_acc, last_tick, last_tempo = 0.0, 0, tempo_events[0][1]
    for tick, tempo in tempo_events:
        t_acc += (tick - last_tick) * last_tempo / 1e6 / division
        seg.append((tick, t_acc, tempo))
        last_tick, last_tempo = tick, tempo

    def t_of(tick: int) -> float:
        k = 0
        for idx in range(len(seg)):
            if seg[idx][0] <= tick:
                k = idx
        tk, ts, tempo = seg[k]
        return ts + (tick - tk) * tempo / 1e6 / division

    seq = NoteSeq()
    for events in tracks:
        open_notes = {}
        for ev in events:
            if ev[1] != "ch":
                continue
            tick, _, hi, ch, d1, d2 = ev
            if hi == 0x90 and d2 > 0:
                open_notes.setdefault((ch, d1), []).append((tick, d2))
            elif hi == 0x80 or (hi == 0x90 and d2 == 0):
                key = (ch, d1)
                if key in open_notes:
                    pending = open_notes[key]
                    close = [(s, v) for s, v in pending if s != tick]
                    keep = [(s, v) for s, v in pending if s == tick]
                    for s, v in close:
                        seq.notes.append(Note(d1, t_of(s), t_of(tick), v, 0))
                    if close and keep:
                        open_notes[key] = keep
                    elif close:
                        del open_notes[key]
            elif hi == 0xB0:
                seq.control_changes.append(ControlChange(t_of(tick), d1, d2))
    seq.total_time = max((n.end_time for n in seq.notes), default=0.0)
    return seq

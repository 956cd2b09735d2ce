"""Reader / writer for serialised ``note_seq.NoteSequence`` protobufs (the ``.proto`` song files of the PIG dataset the
reference loads with ``MidiFile.from_file``: robopianist/music/midi_file.py:175-201) without the ``note_seq`` package.

The protobuf wire format is decoded by hand (varint / 64-bit / length-delimited fields); only what the goal tables
need is kept: notes (pitch, velocity, start_time, end_time, part = fingering), control changes (time, number, value),
``total_time``.  Field numbers are those of note_seq's ``music.proto`` — ``NoteSequence``: notes = 8, total_time = 9,
control_changes = 11, sequence_metadata = 1000 omitted; ``Note``: pitch = 1, velocity = 2, start_time = 3, end_time = 4,
part = 10; ``ControlChange``: time = 1, control_number = 2, control_value = 3 — restated from the published schema
(the package is not installable here: the round-trip below is self-consistent, the numbering itself is UNVERIFIED
against a real PIG file, none of which ships with the reference).
"""
from __future__ import annotations

import struct
from pathlib import Path

from .sequence import ControlChange, Note, NoteSeq

_F_NOTES, _F_TOTAL_TIME, _F_CC = 8, 9, 11


def _varint(buf: bytes, i: int):
    shift = val = 0
    while True:
        b = buf[i]; i += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, i
        shift += 7


def _fields(buf: bytes):
    """Yields (field number, wire type, value) of one message; value = int (varint), float (64-bit as double when asked
    by the caller through the raw bytes) or bytes."""
    i, n = 0, len(buf)
    while i < n:
        key, i = _varint(buf, i)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v = buf[i:i + 8]; i += 8
        elif wt == 2:
            ln, i = _varint(buf, i)
            v = buf[i:i + ln]; i += ln
        elif wt == 5:
            v = buf[i:i + 4]; i += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _signed(v: int) -> int:  # int32 / int64 fields are two's complement varints
    return v - (1 << 64) if v >= 1 << 63 else v


def parse_note_sequence(data: bytes) -> NoteSeq:
    seq = NoteSeq()
    for fno, wt, v in _fields(data):
        if fno == _F_NOTES and wt == 2:
            kw = dict(pitch=0, velocity=0, start_time=0.0, end_time=0.0, part=0)
            for f2, w2, v2 in _fields(v):
                if f2 == 1: kw["pitch"] = _signed(v2)
                elif f2 == 2: kw["velocity"] = _signed(v2)
                elif f2 == 3 and w2 == 1: kw["start_time"] = struct.unpack("<d", v2)[0]
                elif f2 == 4 and w2 == 1: kw["end_time"] = struct.unpack("<d", v2)[0]
                elif f2 == 10: kw["part"] = _signed(v2)
            seq.notes.append(Note(kw["pitch"], kw["start_time"], kw["end_time"], kw["velocity"], kw["part"]))
        elif fno == _F_CC and wt == 2:
            t, num, val = 0.0, 0, 0
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 1: t = struct.unpack("<d", v2)[0]
                elif f2 == 2: num = _signed(v2)
                elif f2 == 3: val = _signed(v2)
            seq.control_changes.append(ControlChange(t, num, val))
        elif fno == _F_TOTAL_TIME and wt == 1:
            seq.total_time = struct.unpack("<d", v)[0]
    return seq


def read_proto(path) -> NoteSeq:
    return parse_note_sequence(Path(path).read_bytes())


# ---- writer (tests, and saving augmented songs the way MidiFile.save does)
def _enc_varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _key(fno, wt):
    return _enc_varint(fno << 3 | wt)


def serialize_note_sequence(seq: NoteSeq) -> bytes:
    out = bytearray()
    for n in seq.notes:
        m = (_key(1, 0) + _enc_varint(n.pitch) + _key(2, 0) + _enc_varint(n.velocity) + _key(3, 1) + struct.pack("<d", n.start_time)
             + _key(4, 1) + struct.pack("<d", n.end_time) + (_key(10, 0) + _enc_varint(n.part) if n.part else b""))
        out += _key(_F_NOTES, 2) + _enc_varint(len(m)) + m
    out += _key(_F_TOTAL_TIME, 1) + struct.pack("<d", seq.total_time)
    for c in seq.control_changes:
        m = _key(1, 1) + struct.pack("<d", c.time) + _key(2, 0) + _enc_varint(c.control_number) + _key(3, 0) + _enc_varint(c.control_value)
        out += _key(_F_CC, 2) + _enc_varint(len(m)) + m
    return bytes(out)

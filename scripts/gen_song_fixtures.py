"""Generates robopianist_b200/music/data/*.json from the reference's Rousseau MIDI files.

Run in the build container (needs /root/reference).  The SMF reader is ours
(robopianist_b200/music/sequence.py); the per-note fingerings restate the FINGERING tables of
robopianist/music/library.py:346-385 (Twinkle) and :413-531 (Nocturne), assigned to the notes
sorted by start time exactly as the reference does (library.py:387-396).
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from robopianist_b200.music.sequence import read_midi  # noqa: E402

REF = Path("/root/reference/robopianist/music/data/rousseau")
OUT = Path(__file__).resolve().parents[1] / "robopianist_b200" / "music" / "data"

TWINKLE = [1, 9, 5, 0, 3, 6, 3, 8, 4, 5, 4, 8, 3, 6, 4, 8, 3, 5, 3, 6, 5, 2, 6, 2, 1, 8, 2, 1, 2, 1,
           6, 2, 0, 9]
# pitch names from the reference's comments, used to check our note ordering
TWINKLE_NAMES = ["C5", "C3", "C4", "C5", "G5", "E4", "G5", "C4", "A5", "F4", "A5", "C4", "G5", "E4",
                 "G5", "C4", "F5", "D4", "F5", "B3", "C4", "E5", "A3", "E5", "D5", "F3", "E5", "D5",
                 "E5", "D5", "G3", "E5", "C5", "C3"]
NOCTURNE = [0, 8, 4, 9, 6, 8, 6, 5, 9, 9, 6, 2, 8, 6, 5, 3, 8, 2, 9, 6, 8, 6, 5, 7, 1, 9, 6, 8, 6, 5,
            0, 7, 4, 9, 6, 8, 6, 5, 0, 3, 2, 1, 0, 7, 4, 9, 6, 8, 6, 5, 0, 7, 2, 9, 6, 7, 6, 5, 7, 2,
            9, 6, 7, 6, 5, 2, 7, 1, 9, 6, 8, 6, 5, 9, 4, 9, 5, 6, 5, 5, 0, 9, 2, 9, 6, 7, 6, 5, 9, 1,
            9, 6, 7, 6, 5, 8, 0, 9, 6, 4, 8, 6, 5, 3, 8, 2, 1, 9, 6, 0, 3, 8, 5, 0, 1, 7, 2]
NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def name_of(p):
    return NAMES[p % 12] + str(p // 12 - 1)


def dump(mid, fingering, title, out, names=None):
    seq = read_midi(REF / mid)
    order = sorted(range(len(seq.notes)), key=lambda i: seq.notes[i].start_time)
    assert len(order) == len(fingering), (len(order), len(fingering))
    for rank, i in enumerate(order):
        seq.notes[i].part = fingering[rank]
        if names is not None:
            assert name_of(seq.notes[i].pitch) == names[rank], (rank, name_of(seq.notes[i].pitch), names[rank])
    d = dict(title=title, artist="Rousseau", total_time=seq.total_time,
             notes=[[n.pitch, n.start_time, n.end_time, n.velocity, n.part] for n in seq.notes],
             control_changes=[[c.time, c.control_number, c.control_value] for c in seq.control_changes])
    (OUT / out).write_text(json.dumps(d))
    print(out, len(seq.notes), "notes", len(seq.control_changes), "cc", "total_time", seq.total_time)


dump("twinkle-twinkle-trimmed.mid", TWINKLE, "Twinkle Twinkle (YouTube)", "twinkle_rousseau.json", TWINKLE_NAMES)
dump("nocturne-trimmed.mid", NOCTURNE, "Nocturne (YouTube)", "nocturne_rousseau.json")

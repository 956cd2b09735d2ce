"""Regenerates robopianist_b200/scene/compiled/*.npz from the reference's assets (build container)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from robopianist_b200.scene import cache, compile as C  # noqa: E402

cache._DIR.mkdir(exist_ok=True)
for task, gc in (("piano_with_shadow_hands", False), ("piano_with_shadow_hands", True), ("self_actuated_piano", False)):
    fm = C.compile_scene(task=task, gravity_compensation=gc)
    path = cache._DIR / cache._name(task, gc, 0.005, 0.05)
    fm.save(path)
    print(path.name, path.stat().st_size, "bytes")

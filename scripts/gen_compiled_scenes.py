"""Regenerates robopianist_b200/scene/compiled/*.npz from the reference's assets (build container)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from robopianist_b200.scene import cache, compile as C  # noqa: E402

cache._DIR.mkdir(exist_ok=True)
configs = [("piano_with_shadow_hands", False, {}), ("piano_with_shadow_hands", True, {}), ("self_actuated_piano", False, {}),
           # the reference's RL training configuration (robopianist-rl README: --reduced-action-space --gravity-compensation)
           ("piano_with_shadow_hands", True, dict(reduced_action_space=True)),
           ("piano_with_shadow_hands", False, dict(reduced_action_space=True)),
           # test fixtures of the other model-changing kwargs (tasks/base.py:102-107)
           ("piano_with_shadow_hands", True, dict(attachment_yaw=10.0)),
           ("piano_with_shadow_hands", True, dict(forearm_dofs=("forearm_tx",))),
           ("piano_with_shadow_hands", True, dict(forearm_dofs=(), reduced_action_space=True)),
           # hand geoms may collide with each other (the reference's default: disable_hand_collisions=False)
           ("piano_with_shadow_hands", True, dict(disable_hand_collisions=False)),
           ("piano_with_shadow_hands", False, dict(disable_hand_collisions=False)),
           ("piano_with_shadow_hands", True, dict(disable_hand_collisions=False, reduced_action_space=True))]
for task, gc, extra in configs:
    fm = C.compile_scene(task=task, gravity_compensation=gc, **extra)
    path = cache._DIR / cache._name(task, gc, 0.005, 0.05, **extra)
    fm.save(path)
    print(path.name, path.stat().st_size, "bytes")

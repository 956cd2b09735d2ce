"""Builds tests/golden/twinkle_ctrl.npz from the reference's recorded action replay
(/root/reference/examples/twinkle_twinkle_actions.npy, canonical [-1, 1] actions, (158, 45) float32;
used by tutorial.ipynb cell 18 and examples/piano_with_shadow_hands_env.py:115-124).  The canonical
actions are rescaled to the action spec exactly as dm_env_wrappers.CanonicalSpecWrapper does:
``min + (a + 1) / 2 * (max - min)`` with the actuator ctrlranges of the compiled scene."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from robopianist_b200.scene import cache  # noqa: E402

acts = np.load("/root/reference/examples/twinkle_twinkle_actions.npy").astype(np.float64)
fm = cache.load_scene(gravity_compensation=True)
lo = np.concatenate([fm.actuator_ctrlrange[:, 0], [0.0]])
hi = np.concatenate([fm.actuator_ctrlrange[:, 1], [1.0]])
scaled = lo + (acts + 1.0) / 2.0 * (hi - lo)
out = Path(__file__).resolve().parents[1] / "tests" / "golden" / "twinkle_ctrl.npz"
np.savez_compressed(out, action=scaled)
print(out, scaled.shape, scaled.dtype)

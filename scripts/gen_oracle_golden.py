"""Golden oracle trajectory (regression fixture): qpos after 10/30/60 control steps of the action replay."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
from oracle import oracle as O  # noqa: E402
from robopianist_b200.scene import cache  # noqa: E402

fm = cache.load_scene(gravity_compensation=True)
om = O.OracleModel(fm); om.set_option(maxcon=24)
d = O.OracleData(om); d.reset()
acts = helpers.twinkle_actions()
out = {}
for t in range(60):
    d.control_step(acts[t, :44])
    if t in (9, 29, 59):
        out[f"qpos_{t}"] = d.qpos.copy()
np.savez_compressed(helpers.GOLDEN / "oracle_twinkle_traj.npz", **out)
print("written", list(out))

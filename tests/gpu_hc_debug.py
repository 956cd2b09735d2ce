"""Debug: error growth of the hand-collision rollout against the oracle (not a test)."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers
from oracle import oracle as O
import test_gpu_hand_collisions as T
fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=1, scene_kw=T.HC)
batch.set_block_envs(1)
om = O.OracleModel(fm); om.set_option(maxcon=40)
od = O.OracleData(om); od.reset()
nv, nu = fm.meta["nv"], fm.meta["nu"]
batch.set_state(np.zeros((1, nv)), np.zeros((1, nv)), np.zeros((1, nv)))
od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
rng = np.random.default_rng(5); target = rng.random(nu)
gn = fm.meta["geom_names"]
prev = 0
for t in range(700):
    if t % 100 == 0: target = rng.random(nu)
    if t % 10 == 0:
        ctrl = T._controls(fm, t, target); ct = T._dev(ctrl[None])
    batch.physics_step(ct.data_ptr(), 1); od.mj_step(ctrl)
    q, v, w, _ = batch.get_state()
    e = np.abs(q[0] - od.qpos).max(); ew = np.abs(w[0] - od.qacc_warmstart).max()
    if e > 3 * prev + 1e-14 or t % 50 == 49:
        cs = [(gn[c["geom1"]].split("/")[1][3:], gn[c["geom2"]].split("/")[1][3:] if fm.geom_hand[c["geom2"]] >= 0 else "key", round(c["dist"], 5)) for c in od.contacts()]
        dg = batch.get_diag()
        print(t, "err q %.2e qacc %.2e" % (e, ew), "ncon", od.ncon, dg["ncon"][0], "it", od.solver_niter, dg["niter"][0], cs)
        prev = max(prev, e)

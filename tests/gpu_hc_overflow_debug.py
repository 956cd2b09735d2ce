"""Debug: contact records of the kernel against the oracle's pruned contact list around the first record overflow (not a test)."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers
from oracle import oracle as O
import test_gpu_hand_collisions as T
fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=1, scene_kw=T.HC)
batch.set_block_envs(1); batch.debug_enable(True)
om = O.OracleModel(fm); om.set_option(maxcon=60); om.set_record_cap(21)
od = O.OracleData(om); od.reset()
nv, nu = fm.meta["nv"], fm.meta["nu"]
batch.set_state(np.zeros((1, nv)), np.zeros((1, nv)), np.zeros((1, nv)))
od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
names = fm.meta["actuator_names"]; lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
rng = np.random.default_rng(3); target = rng.random(nu)
gn = fm.meta["geom_names"]; bn = fm.meta["body_names"]
shown = 0
for t in range(700):
    if t % 100 == 0: target = rng.random(nu)
    if t % 10 == 0:
        ctrl = lo + (hi - lo) * target
        for i, n in enumerate(names):
            if n.endswith("forearm_tx"): ctrl[i] = (-0.2 if "rh_" in n else 0.2) * min(1.0, t / 200.0)
            if n.endswith("forearm_ty"): ctrl[i] = 0.0
        ct = T._dev(ctrl[None])
    if t % 25 == 0 and t:
        batch.set_state(od.qpos[None].copy(), od.qvel[None].copy(), od.qacc_warmstart[None].copy())
    batch.physics_step(ct.data_ptr(), 1); od.mj_step(ctrl)
    q, v, w, _ = batch.get_state()
    e = np.abs(q[0] - od.qpos).max()
    if od.con_overflow and shown < 3:
        shown += 1
        d = batch.debug_read()[0]
        nrec = int(d[640]); rec = d[656:656 + 16 * nrec].reshape(nrec, 16)
        print("t", t, "err", e, "oracle ncon", od.ncon, "kernel records", nrec)
        for c in od.contacts():
            print("   oracle", gn[c["geom1"]].split("/")[1], gn[c["geom2"]].split("/")[1] if fm.geom_hand[c["geom2"]] >= 0 else "key%d" % c["key"], "%.6e" % c["dist"])
        for r in rec:
            print("   kernel key %d hand %d link %d f %.4e %.4e %.4e" % (r[0], r[1], r[2], r[3], r[4], r[5]))
        f = od.efc_force if hasattr(od, "efc_force") else None
    elif e > 1e-9 and shown < 6:
        shown += 1; print("t", t, "err", e, "overflow", od.con_overflow, od.ncon)

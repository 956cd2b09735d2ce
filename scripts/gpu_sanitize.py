"""Workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): a few control steps of every kernel
instantiation — default (1 and 3 environments per block), hand collisions with cross-hand, same-hand and piano contacts, the
reduced action space (padding links), contact-record overflow, random hand shifts, MIDI events, the keys-only task — on a handful of environments.
usage: compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
import test_gpu_hand_collisions as T  # noqa: E402


def run(scene_kw, block_envs, nenv, steps, drive_together=False, task="piano_with_shadow_hands", amp=0.6, **kw):
    fm, model, b = helpers.make_native(task, nenv=nenv, scene_kw=scene_kw, **kw)
    if task == "piano_with_shadow_hands":
        b.set_block_envs(block_envs)
        b.randomize_hand_positions(True, seed=3)
        b.midi_enable(256)
    obs = torch.empty((nenv, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(nenv, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(nenv, dtype=torch.uint8, device="cuda")
    b.reset(None, obs.data_ptr())
    rng = np.random.default_rng(0)
    lo = np.append(fm.actuator_ctrlrange[:, 0], 0.0); hi = np.append(fm.actuator_ctrlrange[:, 1], 1.0)
    target = rng.random((nenv, model.act_dim))
    for t in range(steps):
        a = lo + (hi - lo) * (0.5 + amp * (target - 0.5))
        if drive_together:
            for i, n in enumerate(fm.meta["actuator_names"]):
                if n.endswith("forearm_tx"):
                    a[:, i] = (-0.2 if "rh_" in n else 0.2) * min(1.0, t / 20.0)
        b.step(torch.tensor(a, dtype=torch.float64, device="cuda").data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    torch.cuda.synchronize()
    d = b.get_diag()
    print(scene_kw, "block_envs", block_envs, "ncon", d["ncon"].tolist(), "niter", d["niter"].tolist(), "flags", d["flags"].tolist(), flush=True)
    b.close()


def run_overflow(substeps=130):
    """The schedule of test_beyond_the_record_capacity_the_deepest_contacts_are_kept up to its first overflows (substep 115)."""
    fm, model, b = helpers.make_native("piano_with_shadow_hands", nenv=4, scene_kw=T.HC)
    b.set_block_envs(3)
    nv, nu = fm.meta["nv"], fm.meta["nu"]
    b.set_state(np.zeros((4, nv)), np.zeros((4, nv)), np.zeros((4, nv)))
    names = fm.meta["actuator_names"]; lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    rng = np.random.default_rng(3); target = rng.random(nu)
    seen = 0
    for t in range(substeps):
        if t % 100 == 0:
            target = rng.random(nu)
        if t % 10 == 0:
            ctrl = lo + (hi - lo) * target
            for i, n in enumerate(names):
                if n.endswith("forearm_tx"):
                    ctrl[i] = (-0.2 if "rh_" in n else 0.2) * min(1.0, t / 200.0)
                if n.endswith("forearm_ty"):
                    ctrl[i] = 0.0
            ct = torch.tensor(np.tile(ctrl, (4, 1)), dtype=torch.float64, device="cuda")
        b.physics_step(ct.data_ptr(), 1)
        if t >= 110:
            seen += int((b.get_diag()["flags"] & 1).any())
    print("contact-record overflow: substeps with the flag set", seen, flush=True)
    assert seen > 0
    b.close()


steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
run({}, 1, 4, steps)
run({}, 3, 7, steps)
run(dict(disable_hand_collisions=False), 1, 4, steps, drive_together=True)
run(dict(disable_hand_collisions=False), 3, 7, steps, drive_together=True)
run(dict(disable_hand_collisions=False, reduced_action_space=True), 3, 5, steps, drive_together=True)
run_overflow()
run(None, 1, 4, 5, task="self_actuated_piano", gravity_compensation=False, lookahead=0)
print("workload done")

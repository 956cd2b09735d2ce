#!/usr/bin/env python
"""BASELINE.json configs 2, 3 and 5 on one GPU (run on the GPU box): env-steps/s per config and, for the
PGS-iteration sweep of config 5, the L-inf qpos distance to the 100-iteration run after the timed steps.
Prints one JSON line per measurement.  usage: python scripts/gpu_configs.py [steps]
Actions: config 2 uniform in the action spec (seed 12345, as self_actuated_piano_test.py:151-157); configs 3 / 5
the recorded Twinkle replay + N(0, 0.05) noise per env (there is no recorded policy for the C-major scale)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402


def run(task, song, N, steps, iterations=None, warm=3, seed=12345, scene_kw=None):
    fm, model, b = helpers.make_native(task, song=song, nenv=N, scene_kw=scene_kw)
    if iterations is not None:
        b.set_option(iterations=iterations)
    lo = np.concatenate([fm.actuator_ctrlrange[:, 0], [0.0]]); hi = np.concatenate([fm.actuator_ctrlrange[:, 1], [1.0]])
    rng = np.random.default_rng(seed)
    total = warm + steps
    if task == "piano_with_shadow_hands":
        from robopianist_b200.scene import cache
        base = helpers.reduced_actions(cache.load_scene(gravity_compensation=True), fm, helpers.twinkle_actions())
        acts = np.stack([np.clip(base[t % 158] + rng.normal(0, 0.05, (N, model.act_dim)) * (hi - lo) / 2, lo, hi) for t in range(total)])
    else:
        acts = rng.uniform(lo, hi, size=(total, N, model.act_dim))
    a = torch.tensor(acts, dtype=torch.float64, device="cuda")
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda")
    b.reset(None, obs.data_ptr())
    for t in range(warm):
        b.step(a[t].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(warm, total):
        b.step(a[t].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    d = b.get_diag()
    q = b.get_state()[0]
    out = dict(task=task, song=song, envs=N, pgs_iterations=iterations if iterations is not None else int(fm.meta.get("iterations", 50)),
               steps=steps, ms_per_step=ms, env_steps_per_s=N / ms * 1e3, mean_sweeps_last_substep=float(d["niter"].mean()),
               mean_contacts=float(d["ncon"].mean()), max_contacts=int(d["ncon"].max()), flagged_envs=int(((d["flags"] & 7) != 0).sum()),
               scene=dict(scene_kw or {}))
    b.close()
    return out, q


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    r, _ = run("self_actuated_piano", "TwinkleTwinkleRousseau", 4096, steps, iterations=20)
    print(json.dumps(dict(config=2, **r)), flush=True)
    r, _ = run("piano_with_shadow_hands", "TwinkleTwinkleRousseau", 4096, steps, iterations=50)
    print(json.dumps(dict(config=3, **r)), flush=True)
    ref_q = None
    for it in (100, 75, 50, 35, 20, 10):
        r, q = run("piano_with_shadow_hands", "CMajorScaleTwoHands", 8192, steps, iterations=it)
        if ref_q is None:
            ref_q = q
        r["qpos_linf_vs_100_iterations"] = float(np.abs(q - ref_q).max())
        print(json.dumps(dict(config=5, **r)), flush=True)
    # beyond BASELINE.json: the same 4096-env replay with the hand geoms colliding with each other (the reference's default)
    # and in the reference's RL training configuration (reduced action space + hand collisions + gravity compensation)
    for kw in (dict(disable_hand_collisions=False), dict(disable_hand_collisions=False, reduced_action_space=True)):
        r, _ = run("piano_with_shadow_hands", "TwinkleTwinkleRousseau", 4096, steps, iterations=50, scene_kw=kw)
        print(json.dumps(dict(config="3+" + "+".join(sorted(kw)), **r)), flush=True)

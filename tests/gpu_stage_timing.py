"""Per-phase cycle counts of the step kernel (needs a build with -DRP_PROFILE; ROBOPLAY_LIB selects it).
usage: python tests/gpu_stage_timing.py [N] [step] [noise]   -- bench workload, measured at control step `step`."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers
names = ["kinematics", "RNE+CRB+M scatter", "sweep M^-1", "act/smooth/rows/keys", "barrier waits (setup)", "collision", "contact setup",
         "warm start + cost", "pgs: friction", "pgs: key+limit rows", "pgs: contacts tail", "pgs: iteration barrier", "after pgs", "dual->primal",
         "Euler sweep", "integrate+sync", "pgs contact: load+pre+reduce", "pgs contact: post", "pgs contact: commit", "store"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 888
step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
fm, model, b = helpers.make_native("piano_with_shadow_hands", nenv=N)
acts = helpers.twinkle_actions()
lo = np.concatenate([fm.actuator_ctrlrange[:, 0], [0.0]]); hi = np.concatenate([fm.actuator_ctrlrange[:, 1], [1.0]])
rng = np.random.default_rng(1)
obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda"); rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew); st = torch.empty(N, dtype=torch.uint8, device="cuda")
b.reset(None, obs.data_ptr())
def act(t):
    a = np.tile(acts[t], (N, 1)) + rng.normal(0, noise, (N, 45)) * (hi - lo) / 2
    return torch.tensor(np.clip(a, lo, hi), dtype=torch.float64, device="cuda")
for t in range(step):
    b.step(act(t).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
b.debug_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a = act(step)
e0.record()
b.step(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
e1.record()
torch.cuda.synchronize()
d = b.debug_read()
dg = b.get_diag()
NP = 24
print("envs", N, "step", step, "ms", e0.elapsed_time(e1), "niter(last substep) mean", dg["niter"].mean(), "ncon mean", dg["ncon"].mean(), "ncon max", dg["ncon"].max())
for w in range(2):
    t = d[:, 2392 + NP * w: 2392 + NP * w + NP].mean(axis=0) / 10.0
    tot = t.sum()
    print(f"--- warp {w}: total {tot/1e3:.1f} kcycles per env-substep (= {tot/1.965e3:.1f} us at 1.965 GHz)")
    for n, v in zip(names, t):
        if v > 0: print(f"  {n:30s} {v/1e3:9.2f} kcycles {100*v/tot:5.1f}%")

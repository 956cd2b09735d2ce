"""Per-phase cycle counts of the step kernel (needs a build with -DRP_PROFILE; ROBOPLAY_LIB selects it).
usage: python tests/gpu_stage_timing.py [N] [noise]"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers
names = ["kin+RNE+CRB", "M scatter", "sweep M^-1", "act/smooth/rows/keys", "barrier waits (setup)", "collision", "contact setup (warp0)",
         "warm start + cost", "pgs: friction", "pgs: key+limit rows", "pgs: contacts", "pgs: iteration barrier", "(unused)", "dual->primal", "Euler sweep", "integrate+store"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 888
noise = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
fm, model, b = helpers.make_native("piano_with_shadow_hands", nenv=N)
acts = helpers.twinkle_actions()
lo = np.concatenate([fm.actuator_ctrlrange[:, 0], [0.0]]); hi = np.concatenate([fm.actuator_ctrlrange[:, 1], [1.0]])
rng = np.random.default_rng(1)
obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda"); rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew); st = torch.empty(N, dtype=torch.uint8, device="cuda")
b.reset(None, obs.data_ptr())
def act(t):
    a = np.tile(acts[t], (N, 1)) + rng.normal(0, noise, (N, 45)) * (hi - lo) / 2
    return torch.tensor(np.clip(a, lo, hi), dtype=torch.float64, device="cuda")
for t in range(20):
    b.step(act(t).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
b.debug_enable(True)
b.step(act(20).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
torch.cuda.synchronize()
d = b.debug_read()
dg = b.get_diag()
print("envs", N, "niter(last substep) mean", dg["niter"].mean(), "ncon mean", dg["ncon"].mean())
for w in range(2):
    t = d[:, 2392 + 16 * w: 2392 + 16 * w + 16].mean(axis=0)
    tot = t.sum()
    print(f"--- warp {w}: total {tot/1e6:.3f} Mcycles per env-step (= {tot/1.965e6:.3f} ms at 1.965 GHz)")
    for n, v in zip(names, t):
        if v > 0: print(f"  {n:28s} {v/1e3:9.1f} kcycles {100*v/tot:5.1f}%")

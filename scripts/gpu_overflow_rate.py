"""How often the 21-record contact capacity overflows under a uniformly random policy (new action every control step), with the
hand geoms colliding with each other.  usage: python scripts/gpu_overflow_rate.py [envs] [steps]"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for kw in (dict(disable_hand_collisions=False), dict(disable_hand_collisions=False, reduced_action_space=True), {}):
    fm, model, b = helpers.make_native("piano_with_shadow_hands", nenv=N, scene_kw=kw)
    lo = np.append(fm.actuator_ctrlrange[:, 0], 0.0); hi = np.append(fm.actuator_ctrlrange[:, 1], 1.0)
    g = torch.Generator(device="cuda"); g.manual_seed(12345)
    lo_t, hi_t = torch.tensor(lo, device="cuda"), torch.tensor(hi, device="cuda")
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda"); rew = torch.empty(N, dtype=torch.float64, device="cuda")
    dis = torch.empty_like(rew); st = torch.empty(N, dtype=torch.uint8, device="cuda")
    b.reset(None, obs.data_ptr())
    over = np.zeros(N, bool); bad = np.zeros(N, bool); mx = 0; hist = []; nflag = 0
    for t in range(steps):
        a = lo_t + (hi_t - lo_t) * torch.rand((N, model.act_dim), generator=g, device="cuda", dtype=torch.float64)
        b.step(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        d = b.get_diag()
        nflag += int(((d["flags"] & 1) != 0).sum()); over |= (d["flags"] & 1) != 0; bad |= (d["flags"] & 2) != 0; mx = max(mx, int(d["ncon"].max())); hist.append(float(d["ncon"].mean()))
    print(kw, f"envs {N} steps {steps}: control steps with an overflowing substep {100*nflag/(N*steps):.3f} % of {N*steps}, environments that ever overflowed {over.sum()} ({100*over.mean():.2f} %), diverged {bad.sum()}, max contacts {mx}, mean contacts {np.mean(hist):.2f}", flush=True)
    b.close()

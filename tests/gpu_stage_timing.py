"""Per-stage cycle counts of the step kernel (needs a build with EXTRA=-DRP_PROFILE)."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers
names = ["kinematics", "rne/crb/M", "invertM", "act+smooth+rows", "collision", "contact setup", "ckeys+warmstart", "PGS(total incl. below)", "d2p+euler", "pgs:friction", "pgs:ckey+handlim", "pgs:contacts"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1036
fm, model, b = helpers.make_native("piano_with_shadow_hands", nenv=N)
acts = helpers.twinkle_actions()
obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda"); rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew); st = torch.empty(N, dtype=torch.uint8, device="cuda")
b.reset(None, obs.data_ptr())
for t in range(30):
    a = torch.tensor(np.tile(acts[t], (N, 1)), dtype=torch.float64, device="cuda")
    b.step(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
b.debug_enable(True)
a = torch.tensor(np.tile(acts[30], (N, 1)), dtype=torch.float64, device="cuda")
b.step(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
torch.cuda.synchronize()
d = b.debug_read()
t = d[:, 644:656].mean(axis=0)
dg = b.get_diag()
print("niter", dg["niter"].mean(), "ncon", dg["ncon"].mean())
tot = t[:9].sum()
for n, v in zip(names, t):
    print(f"{n:28s} {v/1e3:10.1f} kcycles/env-step  {100*v/tot:5.1f}%")
print("total kcycles", tot / 1e3, "=> ms at 1.965GHz", tot / 1.965e6)

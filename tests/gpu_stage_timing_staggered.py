"""Phase timers (profile build: ROBOPLAY_LIB=.../lib_prof.so) on the full-episode workload of bench.py (episode phases
staggered over the batch): per-stage cycles and the distribution of per-environment durations.
usage: python tests/gpu_stage_timing_staggered.py [N]"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import bench, helpers
names = ["kinematics", "RNE+CRB+M scatter", "sweep M^-1", "act/smooth/rows/keys", "barrier waits (setup)", "collision", "contact setup",
         "warm start + cost", "pgs: friction", "pgs: key+limit rows", "pgs: contacts tail", "pgs: iteration barrier", "after pgs", "dual->primal",
         "Euler sweep", "integrate+sync", "pgs contact: load+pre+reduce", "pgs contact: post", "pgs contact: commit", "store"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
acts_np, fm = bench.make_actions(N, 159, 0)
fm2, model, b = helpers.make_native("piano_with_shadow_hands", nenv=N)
dev = torch.device("cuda")
actions = torch.from_numpy(acts_np).to(dev)
obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device=dev); rew = torch.empty(N, dtype=torch.float64, device=dev); dis = torch.empty_like(rew); st = torch.empty(N, dtype=torch.uint8, device=dev)
s_i = torch.arange(N, device=dev) % 159; ar = torch.arange(N, device=dev)
b.reset(None, obs.data_ptr())
for g in range(159):
    if g > 0:
        b.reset((s_i == g).to(torch.uint8).data_ptr(), obs.data_ptr())
    b.step(actions[(g - s_i) % 159, ar].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
for k in range(3):
    b.step(actions[(159 + k - s_i) % 159, ar].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
b.debug_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a = actions[(162 - s_i) % 159, ar].contiguous()
e0.record(); b.step(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr()); e1.record()
torch.cuda.synchronize()
d = b.debug_read(); dg = b.get_diag()
NP = 24
tot = d[:, 2392:2392 + NP].sum(axis=1)   # warp 0 total cycles per env-step
print("envs", N, "ms", e0.elapsed_time(e1), "ncon mean", dg["ncon"].mean(), "max", dg["ncon"].max(), "niter mean", dg["niter"].mean())
print("per-env duration (Mcycles): mean %.3f  median %.3f  p90 %.3f  max %.3f  min %.3f" % (tot.mean() / 1e6, np.median(tot) / 1e6, np.percentile(tot, 90) / 1e6, tot.max() / 1e6, tot.min() / 1e6))
print("sum / 888 slots = %.3f ms ; max env = %.3f ms" % (tot.sum() / 888 / 1.965e6, tot.max() / 1.965e6))
hist = np.bincount(dg["ncon"], minlength=22); print("ncon histogram", hist.tolist())
for w in range(2):
    t = d[:, 2392 + NP * w: 2392 + NP * w + NP].mean(axis=0) / 10.0
    print(f"--- warp {w}: total {t.sum()/1e3:.1f} kcycles per env-substep")
    for n, v in zip(names, t):
        if v > 0: print(f"  {n:30s} {v/1e3:9.2f} kcycles {100*v/t.sum():5.1f}%")

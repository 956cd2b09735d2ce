"""Timing sweep over batch sizes (GPU box): python tests/gpu_timing.py [N ...]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers  # noqa: E402


def run(N, steps=20, warm=5, task="piano_with_shadow_hands"):
    fm, model, b = helpers.make_native(task, nenv=N)
    acts = helpers.twinkle_actions()
    if task != "piano_with_shadow_hands":
        rng = np.random.default_rng(0)
        acts = rng.uniform(0, 0.07, size=(158, model.act_dim))
    a = torch.tensor(np.tile(acts[:, None, :], (1, N, 1)), dtype=torch.float64, device="cuda")
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda")
    b.reset(None, obs.data_ptr())
    for t in range(warm):
        b.step(a[t].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(warm, warm + steps):
        b.step(a[t].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    d = b.get_diag()
    print(f"{task} N={N}: {ms:.3f} ms/step -> {N / ms * 1e3:.0f} env-steps/s  niter mean {d['niter'].mean():.1f} ncon mean {d['ncon'].mean():.2f}", flush=True)
    b.close()


if __name__ == "__main__":
    Ns = [int(x) for x in sys.argv[1:]] or [37, 148, 296, 592, 1036, 4096]
    for N in Ns:
        run(N)
    run(4096, task="self_actuated_piano")

"""Distribution of per-env step durations in the bench workload (needs a -DRP_PROFILE build via ROBOPLAY_LIB)."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers, bench
N = 4096
actions_np, fm = bench.make_actions(N, 24, 0)
goal, fing = helpers.song_tables(bench.SONG)
from robopianist_b200 import native
model = native.Model(fm, goal, fing, kind=0, n_steps_lookahead=1)
b = native.Batch(model, N)
obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda"); rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew); st = torch.empty(N, dtype=torch.uint8, device="cuda")
acts = torch.from_numpy(actions_np).cuda()
b.reset(None, obs.data_ptr())
b.debug_enable(True)
for t in range(23):
    b.step(acts[t % 24].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
    if t in (5, 12, 22):
        torch.cuda.synchronize()
        d = b.debug_read()
        tot = d[:, 2392:2408].sum(axis=1) / 1.965e6  # ms, warp 0
        dg = b.get_diag()
        q = np.percentile(tot, [0, 10, 50, 90, 99, 100])
        ideal = tot.sum() / 888
        srt = np.sort(tot)[::-1]
        # list-scheduling makespan estimates: natural order vs longest-first
        def makespan(order):
            import heapq
            h = [0.0] * 888
            heapq.heapify(h)
            for x in order: heapq.heappush(h, heapq.heappop(h) + x)
            return max(h)
        print(f"t={t}: per-env ms min/p10/p50/p90/p99/max = {np.round(q, 2)}; sum/888 = {ideal:.2f} ms; makespan natural order {makespan(tot):.2f} ms, longest-first {makespan(srt):.2f} ms; corr(duration, niter_last*..)",
              np.round(np.corrcoef(tot, dg['niter'] * (1 + dg['ncon']))[0, 1], 3))

#!/usr/bin/env python
"""bench.py — env-steps/s of the batched RoboPianist rollout (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels via the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU path, all host cores

One "step" = one control step (10 physics substeps of 5 ms + task logic) of every environment in
the batch.  Workload (config.workload): piano_with_shadow_hands, TwinkleTwinkleRousseau
(trim_silence=True, gravity_compensation=True: the configuration of the reference's recorded action
replay), 4096 envs per GPU, PGS 50 iterations, elliptic cones, fp64.  Env 0 replays the recorded
actions; env i>0 adds N(0, 0.05) noise (canonical action units, seed 12345+i).  Episodes auto-reset.

Prints ONE JSON line (see the task contract): value = whole-job env-steps/s with inputs resident in
HBM; e2e = same metric through rp_step_host with pinned-host actions in / results out every step;
roofline = algorithmic HBM bytes of the step kernel / its CUDA-event duration vs the measured copy
peak; cpu_baseline = the CPU oracle (port of the reference's MuJoCo path; MuJoCo itself is not
installable here) on a bounded sample on this box's cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

ENVS_PER_GPU = 4096
NCU_DRAM_BYTES_PER_LAUNCH = 57.0e6  # profiles/r01_ncu_r01g.csv: dram__bytes_read.sum + dram__bytes_write.sum, 4096 envs
SONG = "TwinkleTwinkleRousseau"
METRIC = "env-steps/sec at 4096 envs/GPU (piano_with_shadow_hands, 10 substeps/step)"


def algorithmic_bytes_per_env_step(nv, nu, obs_dim, act_dim):
    """SURVEY.md §8(d) restated for fp64 storage: state read + write, ctrl, action, obs, scalars."""
    state = 3 * nv * 8
    return 2 * state + act_dim * 8 + 2 * nu * 8 + obs_dim * 8 + (8 + 8 + 1) + 2 * 4 + 8


def make_actions(n_env, n_pool, rank):
    """[n_pool, n_env, 45] float64 actions: recorded replay + per-env noise, clipped to the spec."""
    import helpers
    from robopianist_b200.scene import cache

    fm = cache.load_scene(gravity_compensation=True)
    lo = np.concatenate([fm.actuator_ctrlrange[:, 0], [0.0]])
    hi = np.concatenate([fm.actuator_ctrlrange[:, 1], [1.0]])
    base = helpers.twinkle_actions()  # [158, 45] in spec units
    T = base.shape[0]
    out = np.empty((n_pool, n_env, 45))
    for t in range(n_pool):
        out[t] = base[min(t % (T + 1), T - 1)]
    rng = np.random.default_rng(12345 + 1000003 * rank)
    noise = rng.normal(0.0, 0.05, size=out.shape) * (hi - lo) / 2.0
    noise[:, 0, :] = 0.0 if rank == 0 else noise[:, 0, :]
    return np.clip(out + noise, lo, hi), fm


class ClockSampler:
    def __init__(self, device_index):
        self.samples, self._stop = [], threading.Event()
        self.idx = device_index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append([x.strip() for x in o])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.thread.join(timeout=5)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) >= 6 and s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_run(steps, warmup, sample_envs=None):
    """Times the CPU path (oracle port, all host threads) on a bounded sample of the same workload."""
    import helpers
    from oracle import oracle as O
    from robopianist_b200.scene import cache

    cores = os.cpu_count() or 1
    fm = cache.load_scene(gravity_compensation=True)
    om = O.OracleModel(fm)
    om.set_option(maxcon=24)
    acts = helpers.twinkle_actions()[:, :44]
    E = sample_envs or 4 * cores
    rng = np.random.default_rng(12345)
    total = warmup + steps
    ctrl = np.empty((total, E, 44))
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    for t in range(total):
        ctrl[t] = np.clip(acts[t % 158] + rng.normal(0, 0.05, (E, 44)) * (hi - lo) / 2, lo, hi)
    ctrl[:, 0] = acts[np.arange(total) % 158]
    # the rollout entry restarts from reset; time (warmup+steps) minus warmup-only to isolate `steps`
    t0 = time.perf_counter()
    om.rollout(ctrl[:warmup], E, cores) if warmup else None
    t1 = time.perf_counter()
    om.rollout(ctrl, E, cores)
    t2 = time.perf_counter()
    dt = max((t2 - t1) - (t1 - t0), 1e-9)
    return dict(value=E * steps / dt, cores=cores, envs=E, seconds=dt, ms_per_step=dt / steps * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"piano_with_shadow_hands/{SONG} trim_silence gravcomp primitive_fingertips "
                          f"disable_hand_collisions, {args.envs_per_gpu} envs/GPU, PGS<=50 it tol 1e-8, elliptic, dt=0.005x10",
              "global_envs": args.envs_per_gpu * max(world, 1), "parallelism": f"env-shard x{max(world, 1)}",
              "l2": "flushed between timed steps (256 MiB device memset)",
              "launch_shape": "3 environments per thread block (2 blocks per SM), phase-locked per substep (chosen by batch size; rp_set_block_envs)"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.steps, max(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "env-steps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config, sample=f"{r['envs']} envs x {args.steps} control steps"),
                "cpu_baseline": {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port",
                                 "sample": f"{r['envs']} envs x {args.steps} control steps, {r['cores']} threads"},
                "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from robopianist_b200 import native
    import helpers

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N = args.envs_per_gpu
    pool = max(8, min(64, args.steps + args.warmup))
    actions_np, fm = make_actions(N, pool, rank)
    goal, fing = helpers.song_tables(SONG)
    model = native.Model(fm, goal, fing, kind=0, n_steps_lookahead=1)
    batch = native.Batch(model, N, device=local_rank)
    dev = torch.device("cuda", local_rank)
    actions = torch.from_numpy(actions_np).to(dev)
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device=dev)
    rew = torch.empty(N, dtype=torch.float64, device=dev)
    dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device=dev)
    obs_all = torch.empty((N * world, model.obs_dim), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def one_step(t):
        batch.step(actions[t % pool].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(obs_all, obs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch.reset(None, obs.data_ptr())
    for t in range(args.warmup):
        one_step(t)
    barrier()
    launches0 = batch.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        for k in range(args.steps):
            flush.zero_()
            evs[k][0].record()
            one_step(args.warmup + k)
            evs[k][1].record()
        barrier()
        wall = time.perf_counter() - wall0
    launches = batch.launches - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms.item())
    value = N * world * args.steps / (dev_ms * 1e-3)
    diag = batch.get_diag()

    # ---- e2e: the same step sequence again, through the host call (rp_step_staged: pageable numpy actions ->
    # pinned staging -> device, launch, device -> pinned staging, read on the host), wall clock
    batch.reset(None, obs.data_ptr())
    for t in range(args.warmup):
        one_step(t)
    barrier()
    e2e_steps = args.steps
    host_actions = actions_np
    w0 = time.perf_counter()
    e2e_reward = 0.0
    for k in range(e2e_steps):
        _, rew_h, _, _, _ = batch.step_host(host_actions[(args.warmup + k) % pool], copy=False)
        e2e_reward += float(rew_h.sum())  # each rank consumes its own shard on the host (no collective on this path)
    barrier()
    e2e_s = time.perf_counter() - w0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = N * world * e2e_steps / float(t_e.item())
    h2d = N * model.act_dim * 8
    d2h = N * (model.obs_dim + 2 + 5) * 8 + N

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the step kernel (the only kernel in the step)
    abytes = algorithmic_bytes_per_env_step(model.nv, model.nu, model.obs_dim, model.act_dim)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kernel_ms = dev_ms / args.steps  # one launch per step (plus NCCL for N>1)
    achieved = N * abytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH if N == ENVS_PER_GPU else None, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/)",
                "peak_source": "measured" if peaks else "fallback",
                "algorithmic_bytes_per_env_step": abytes,
                "note": "latency/ALU-bound fp64 Gauss-Seidel; see profiles/ for pipe utilisation"}
    line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks.summary(), "wall_s": wall,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "roofline": roofline,
            "solver": {"mean_pgs_iters_last_substep": float(diag["niter"].mean()),
                       "mean_contacts": float(diag["ncon"].mean()), "flagged_envs": int((diag["flags"] != 0).sum())}}
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(steps=20, warmup=2)
        line["cpu_baseline"] = {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port",
                                "sample": f"{r['envs']} envs x 20 control steps, {r['cores']} threads, {r['seconds']:.1f} s"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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
NCU_DRAM_BYTES_PER_LAUNCH = 48.2e6  # profiles/r02_ncu_end.csv: dram__bytes_read.sum (24.1 MB) + dram__bytes_write.sum (24.1 MB) of one captured launch, 4096 envs
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


def physical_cores():
    """(physical cores, logical CPUs) usable by this process (Linux: unique (package, core) pairs of the affinity set)."""
    try:
        logical = sorted(os.sched_getaffinity(0))
    except Exception:
        logical = list(range(os.cpu_count() or 1))
    pairs = set()
    try:
        for c in logical:
            base = Path(f"/sys/devices/system/cpu/cpu{c}/topology")
            pairs.add((base.joinpath("physical_package_id").read_text().strip(), base.joinpath("core_id").read_text().strip()))
    except Exception:
        pairs = set()
    return (len(pairs) if pairs else len(logical)), len(logical)


def cpu_reference_run(steps, warmup, envs=1024, repeats=3, single_core=True):
    """Times the CPU path (the oracle port of the reference's mj_step path; MuJoCo itself is not installable here) on a
    bounded sample of the bench workload: the first `envs` environments of rank 0's batch, same actions, `warmup`
    untimed control steps from reset, then `repeats` consecutive timed segments of `steps` control steps (the rollout
    simply continues), median.  One persistent workspace per thread (oracle.OraclePool): nothing is allocated inside
    a timed call.  Threads = physical cores.  Also the single-thread figure on a smaller sample."""
    from oracle import oracle as O

    phys, logical = physical_cores()
    nthreads = max(1, min(phys, 256))
    total = warmup + repeats * steps
    ctrl, fm = make_actions(envs, total, 0)
    ctrl = np.ascontiguousarray(ctrl[:, :, :44])
    om = O.OracleModel(fm)
    om.set_option(maxcon=21)
    pool = O.OraclePool(om, envs, nthreads)
    pool.run(ctrl[:max(warmup, 1)])
    times, macs = [], 0
    for r in range(repeats):
        seg = ctrl[max(warmup, 1) + r * steps: max(warmup, 1) + (r + 1) * steps]
        if seg.shape[0] == 0:
            break
        t0 = time.perf_counter()
        macs += pool.run(seg)
        times.append((time.perf_counter() - t0) / seg.shape[0])
    dt = float(np.median(times))
    n_timed = sum(1 for _ in times) * steps * envs
    out = dict(value=envs / dt, cores=nthreads, logical_cpus=logical, envs=envs, seconds=float(np.sum(times) * steps),
               ms_per_step=dt * 1e3, repeats=len(times), spread=float(max(times) / min(times)) if times else None,
               flops_per_env_step=2.0 * macs / max(n_timed, 1))
    pool.close()
    if single_core:
        e1 = 32
        p1 = O.OraclePool(om, e1, 1)
        p1.run(ctrl[:max(warmup, 1), :e1])
        k1 = min(steps, 3)
        t0 = time.perf_counter()
        p1.run(ctrl[max(warmup, 1):max(warmup, 1) + k1, :e1])
        out["single_core_value"] = e1 * k1 / (time.perf_counter() - t0)
        p1.close()
    return out


FP64_PEAK_TFLOPS = 148 * 1.8 * 32 * 2 * 1.965e9 / 1e12  # measured DFMA issue rate 1.8 warp-instr/clk/SM (profiles/r01_ubench3_pipe_throughput.txt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--total-envs", type=int, default=0, help="strong scaling: fixed global batch split over the ranks (BASELINE config 4: 16384)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-episode", action="store_true", help="skip the second measurement with episode phases staggered across the batch")
    ap.add_argument("--no-hand-collisions", action="store_true", help="skip the extra measurement with the hand geoms colliding with each other")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: whatever else a library writes there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, ...) is sent to stderr; the line itself goes out through the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    strong = args.total_envs > 0
    if strong:
        if args.total_envs % max(world, 1):
            raise SystemExit("--total-envs must be a multiple of the number of ranks")
        args.envs_per_gpu = args.total_envs // max(world, 1)
    N = args.envs_per_gpu
    config = {"workload": f"piano_with_shadow_hands/{SONG} trim_silence gravcomp primitive_fingertips "
                          f"disable_hand_collisions, {N} envs/GPU, PGS<=50 it tol 1e-8, elliptic, dt=0.005x10; "
                          "40 box/cylinder/mesh hand colliders of the reference model are outside the narrowphase's "
                          "coverage (none within reach of the piano: solver.flagged_envs)",
              "global_envs": N * max(world, 1), "parallelism": f"env-shard x{max(world, 1)}",
              "l2": "flushed between timed steps (256 MiB device memset)",
              "launch_shape": "3 environments per thread block (2 blocks per SM), phase-locked per substep (chosen by batch size; rp_set_block_envs)"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.steps, max(args.warmup, 1))
        sample = (f"first {r['envs']} envs of the batch x {args.steps} control steps x {r['repeats']} consecutive segments (median) after "
                  f"{max(args.warmup, 1)} warm-up steps, {r['cores']} threads (physical cores; {r['logical_cpus']} logical CPUs)")
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "env-steps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config, sample=sample),
                "cpu_baseline": {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port", "sample": sample,
                                 "single_core_value": r.get("single_core_value"), "segment_spread": r["spread"],
                                 "note": "CPU restatement of the reference's MuJoCo path (MuJoCo itself is not installable here and would be several times faster per core)"},
                "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    import warnings
    from robopianist_b200.music import library
    from robopianist_b200.suite.sharded import ShardedPianoEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    T_CYCLE = 159                    # 158 control steps + the auto-reset step
    pool = T_CYCLE if not args.no_full_episode else max(8, min(64, args.steps + args.warmup))
    actions_np, fm = make_actions(N, pool, rank)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)   # the uncovered-collider notice: part of config.workload here
        env = ShardedPianoEnv(library.load(SONG), N * world, device=local_rank, trim_silence=True, gravity_compensation=True,
                              primitive_fingertip_collisions=True, disable_hand_collisions=True, n_steps_lookahead=1)
    loc = env.local                  # BatchedPianoEnv of this rank
    model, batch = loc.model, loc.batch
    actions = torch.from_numpy(actions_np).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(step_fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream (the L2 flush stays outside the brackets); the
        gather of step k overlaps step k+1 — a wait for it lands inside a later bracket, the last one inside its own."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush.zero_()
            evs[k][0].record()
            step_fn(k)
            if k == steps - 1:
                env.gathered()
            evs[k][1].record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- protocol A: reset, W warm-up steps, K timed steps (the window the driver uses)
    env.reset()
    for t in range(args.warmup):
        env.step(actions[t % pool])
    env.gathered()
    barrier()
    launches0 = batch.launches
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        dev_ms = timed_region(lambda k: env.step(actions[(args.warmup + k) % pool]), args.steps)
        wall = time.perf_counter() - wall0
    launches = batch.launches - launches0
    value = N * world * args.steps / (dev_ms * 1e-3)
    diag = batch.get_diag()

    # ---- e2e: the same step sequence again through the host call of the public API (BatchedPianoEnv.step_host: pageable
    # numpy actions -> pinned staging -> device, launch, device -> pinned staging, read on the host), wall clock
    loc.reset()
    for t in range(args.warmup):
        loc.step(actions[t % pool])
    barrier()
    w0 = time.perf_counter()
    e2e_reward = 0.0
    for k in range(args.steps):
        _, rew_h, _, _, _ = loc.step_host(actions_np[(args.warmup + k) % pool], copy=False)
        e2e_reward += float(rew_h.sum())  # each rank consumes its own shard on the host (no collective on this path)
    barrier()
    e2e_s = time.perf_counter() - w0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = N * world * args.steps / float(t_e.item())
    h2d = N * model.act_dim * 8
    d2h = N * (model.obs_dim + 2 + 5) * 8 + N

    # ---- protocol B: episode phases staggered uniformly over the batch (env i is reset at global step i mod 159 of a 159
    # step lead-in), so that the timed window holds every part of the song in equal measure: the full-episode mean
    full = None
    if not args.no_full_episode:
        s_i = torch.arange(N, device=dev) % T_CYCLE
        ar = torch.arange(N, device=dev)
        loc.reset()
        for g in range(T_CYCLE):
            m = (s_i == g).to(torch.uint8)
            if g > 0:
                loc.reset(m)
            loc.step(actions[(g - s_i) % T_CYCLE, ar])
        staged = [actions[(T_CYCLE + k - s_i) % T_CYCLE, ar].contiguous() for k in range(args.steps)]
        barrier()
        ms_b = timed_region(lambda k: env.step(staged[k]), args.steps)
        dg = batch.get_diag()
        full = {"ms_per_step": ms_b / args.steps, "value": N * world * args.steps / (ms_b * 1e-3), "steps": args.steps,
                "mean_contacts": float(dg["ncon"].mean()), "mean_pgs_iters_last_substep": float(dg["niter"].mean()),
                "flagged_envs": int((dg["flags"] != 0).sum()),
                "how": "episode phases staggered uniformly (env i reset at lead-in step i mod 159): every timed step sees the whole song's contact mix"}

    # ---- protocol C: the same window with the hand geoms colliding with each other (disable_hand_collisions=False, the
    # reference's default for that kwarg; SURVEY.md 8(d) config 3 leaves it at the default): a second batch on the same
    # device, same actions, same protocol.  Reported next to the headline, which stays on north_star's configuration.
    hc = None
    if not args.no_hand_collisions:
        from robopianist_b200.suite.env import BatchedPianoEnv
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            env_hc = BatchedPianoEnv(library.load(SONG), num_envs=N, device=local_rank, trim_silence=True, gravity_compensation=True,
                                     primitive_fingertip_collisions=True, disable_hand_collisions=False, n_steps_lookahead=1)
        env_hc.reset()
        for t in range(args.warmup):
            env_hc.step(actions[t % pool])
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            flush.zero_()
            evs[k][0].record()
            env_hc.step(actions[(args.warmup + k) % pool])
            evs[k][1].record()
        barrier()
        t_hc = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_hc, op=dist.ReduceOp.MAX)
        dg = env_hc.batch.get_diag()
        hc = {"ms_per_step": float(t_hc.item()) / args.steps, "value": N * world * args.steps / (float(t_hc.item()) * 1e-3), "steps": args.steps,
              "mean_contacts": float(dg["ncon"].mean()), "mean_pgs_iters_last_substep": float(dg["niter"].mean()),
              "flagged_envs": int(((dg["flags"] & 7) != 0).sum()),
              "workload": "as config.workload with disable_hand_collisions=False: the finger / thumb spheres and capsules of both hands also "
                          "collide with each other (596 pairs); kernel instantiation rp_step_kernel<2, 3, true>"}
        env_hc.close()

    if rank != 0:
        env.close()
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the step kernel (the only kernel of a step; N>1 adds the all-gather)
    abytes = algorithmic_bytes_per_env_step(model.nv, model.nu, model.obs_dim, model.act_dim)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kernel_ms = dev_ms / args.steps
    achieved = N * abytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH if N == ENVS_PER_GPU else None,
                "traffic_unit": "bytes/launch: ncu dram__bytes_read.sum + dram__bytes_write.sum of one captured launch, cited from profiles/ (NOT measured in this run)",
                "peak_source": "measured" if peaks else "fallback",
                "algorithmic_bytes_per_env_step": abytes,
                "note": "latency-bound fp64 Gauss-Seidel: the HBM fraction is reported as required, the fp64 figures below say how far the solver is from the arithmetic roof"}
    line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks.summary(), "wall_s": wall,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.steps, "api": "BatchedPianoEnv.step_host"},
            "gpu_launches": int(launches), "roofline": roofline,
            "solver": {"mean_pgs_iters_last_substep": float(diag["niter"].mean()),
                       "mean_contacts": float(diag["ncon"].mean()), "flagged_envs": int((diag["flags"] != 0).sum())}}
    if full is not None:
        line["full_episode"] = full
    if hc is not None:
        line["hand_collisions"] = hc
    if world > 1:
        line["collective"] = {"op": "all_gather_into_tensor (NCCL), async, overlapped with the next step's kernel", "dtype": "float32",
                              "bytes_received_per_rank_per_step": env.gather_bytes_per_step}
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(steps=5, warmup=2, envs=512)
        line["cpu_baseline"] = {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port",
                                "sample": f"first {r['envs']} envs x 5 control steps x {r['repeats']} consecutive segments (median) after 2 warm-up steps, "
                                          f"{r['cores']} threads (physical cores; {r['logical_cpus']} logical CPUs), {r['seconds']:.1f} s",
                                "single_core_value": r.get("single_core_value"), "segment_spread": r["spread"]}
        # solver arithmetic: multiply-adds the oracle counts for the same workload (dense A+R formulation: projection
        # J M^-1 J^T and the Gauss-Seidel residuals), two flops each
        fl = r["flops_per_env_step"]
        roofline.update({"flops_per_env_step": fl, "flops_source": "CPU oracle MAC counter on the same workload (first control steps of the sample), x2",
                         "achieved_tflops": fl * N / (kernel_ms * 1e-3) / 1e12, "fp64_peak_tflops": FP64_PEAK_TFLOPS,
                         "fp64_peak_source": "DFMA issue rate measured on B200 (scratch/ubench3.cu: 1.8 warp-instr/clk/SM x 148 SMs x 1.965 GHz)",
                         "fp64_frac": fl * N / (kernel_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                         "arithmetic_intensity_flop_per_byte": fl / abytes})
    emit(line)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

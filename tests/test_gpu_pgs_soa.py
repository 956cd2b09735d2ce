"""EXPERIMENTAL kernel of the split-solver groundwork (csrc/rp_pgs_kernel.cu, lane == environment) against the
host build of the same solver and the oracle, on problems rebuilt from the oracle; prints its time for a
4096-environment batch of such problems (a measurement for DESIGN.md §5, not an assertion)."""
import ctypes as C

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from robopianist_b200 import native
from robopianist_b200.scene import cache
from test_pgs_soa import build_problem, hd, read_forces  # noqa: F401  (hd: fixture)


@pytest.mark.gpu
def test_soa_kernel_matches_host_solver_and_oracle(hd):
    lib, off, size = hd
    fm = cache.load_scene(gravity_compensation=True)
    om = O.OracleModel(fm)
    om.set_option(maxcon=24)
    acts = helpers.twinkle_actions()[:, :44]
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    rng = np.random.default_rng(3)
    E = 8
    drivers = [O.OracleData(om) for _ in range(E)]
    probe = O.OracleData(om)
    for d in drivers:
        d.reset()
    scale = 1.0 / (fm.meta["meaninertia"] * fm.meta["nv"])
    it, tol = fm.meta["iterations"], fm.meta["tolerance"]
    P = maps = finals = niters = None
    for t in range(24):
        ctrls = [np.clip(acts[t] + (rng.normal(0, 0.05, 44) * (hi - lo) / 2 if e else 0), lo, hi) for e in range(E)]
        if t == 23:
            P = np.zeros((size, E)); maps, finals, niters = [], [], []
            for e, d in enumerate(drivers):
                probe.qpos[:] = d.qpos; probe.qvel[:] = d.qvel; probe.qacc_warmstart[:] = d.qacc_warmstart
                probe.ctrl[:] = ctrls[e]
                probe.step1(); probe.call("orc_smooth"); probe.call("orc_solve_init")
                P[:, e], rm = build_problem(fm, probe, off, size)
                maps.append(rm)
                probe.call("orc_solve_iterate")
                finals.append(probe.efc("efc_force")); niters.append(probe.solver_niter)
        for e, d in enumerate(drivers):
            d.control_step(ctrls[e])
    assert sum(int(P[0, e]) for e in range(E)) > 0  # contacts present
    host = np.ascontiguousarray(P.copy())
    for e in range(E):
        lib.hd_pgs_solve(host.ctypes.data_as(C.POINTER(C.c_double)), E, e, it, tol, scale)
    N = 4096
    for staged in (False, True):
        dev, _ = native.debug_pgs_solve(P.copy(), it, tol, scale, staged=staged)
        for e in range(E):
            got = read_forces(dev[:, e], off, maps[e], len(finals[e]))
            ref = read_forces(host[:, e], off, maps[e], len(finals[e]))
            assert int(dev[3, e]) == int(host[3, e]) == niters[e]
            assert np.abs(got - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
            assert np.abs(got - finals[e]).max() < 1e-9 * max(1.0, np.abs(finals[e]).max())
        # timing: the same problems tiled to the bench batch size
        big = np.ascontiguousarray(np.tile(P, (1, N // E)))
        _, ms = native.debug_pgs_solve(big, it, tol, scale, repeat=5, staged=staged)
        print(f"\nSoA Gauss-Seidel kernel ({'scalar rows staged in shared memory' if staged else 'all rows from global memory'}), {N} environments, "
              f"mean contacts {P[0].mean():.2f}, mean sweeps {np.mean(niters):.1f}: {ms:.3f} ms per substep solve")

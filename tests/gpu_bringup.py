"""GPU bring-up: stage-by-stage comparison of the CUDA substep against the CPU oracle.
Run on the GPU box:  python tests/gpu_bringup.py  (prints a report; not collected by pytest)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers  # noqa: E402
from oracle import oracle as O  # noqa: E402

np.set_printoptions(linewidth=220, precision=6, suppress=False)
D = dict(QS=0, QA=160, QC=320, BIAS=480, SC=640, CON=656, M=1040, MINV=1716)


def err(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) if len(np.atleast_1d(a)) else 0.0


def compare_substep(tag, fm, batch, od, q, v, w, ctrl):
    nv = fm.meta["nv"]
    nh = fm.meta["nhand"]
    N = batch.nenv
    batch.set_state(np.tile(q, (N, 1)), np.tile(v, (N, 1)), np.tile(w, (N, 1)))
    ct = torch.tensor(np.tile(ctrl, (N, 1)), dtype=torch.float64, device="cuda")
    batch.debug_enable(True)
    batch.physics_step(ct.data_ptr(), 1)
    torch.cuda.synchronize()
    dbg = batch.debug_read()[0]
    q2, v2, w2, _ = batch.get_state()
    od.qpos[:] = q; od.qvel[:] = v; od.qacc_warmstart[:] = w; od.ctrl[:] = ctrl
    od.step1()
    M = od.M.copy()
    od.step2()
    print(f"== {tag}: oracle ncon={od.ncon} nefc={od.nefc} niter={od.solver_niter} | cuda ncon={dbg[D['SC']]:.0f} niter={dbg[D['SC']+1]:.0f} ovf={dbg[D['SC']+2]:.0f} wscost={dbg[D['SC']+3]:.3e}")
    if nh:
        Mh = M[88:114, 88:114]
        print("   M hand0      ", err(dbg[D['M']:D['M'] + 676].reshape(26, 26), Mh))
        print("   Minv hand0   ", err(dbg[D['MINV']:D['MINV'] + 676].reshape(26, 26), np.linalg.inv(Mh)), "scale", np.abs(np.linalg.inv(Mh)).max())
        print("   bias         ", err(dbg[D['BIAS'] + 88:D['BIAS'] + nv], od.qfrc_bias[88:] - 0 * od.qfrc_passive[88:]),
              "(oracle bias excludes gravcomp; cuda folds gravcomp into bias)")
    print("   qacc_smooth  ", err(dbg[D['QS']:D['QS'] + nv], od.qacc_smooth), "scale", np.abs(od.qacc_smooth).max())
    print("   qfrc_constr  ", err(dbg[D['QC']:D['QC'] + nv], od.qfrc_constraint), "scale", np.abs(od.qfrc_constraint).max())
    print("   qacc         ", err(dbg[D['QA']:D['QA'] + nv], od.qacc), "scale", np.abs(od.qacc).max())
    print("   qpos'        ", err(q2[0], od.qpos), " qvel' ", err(v2[0], od.qvel), " ws' ", err(w2[0], od.qacc_warmstart))
    nc = int(dbg[D['SC']])
    cons = od.contacts()
    for c in range(min(nc, len(cons), 4)):
        o = dbg[D['CON'] + 16 * c: D['CON'] + 16 * c + 16]
        adr = 52 * (1 if nh else 0) + int((od.efc_type() == 1).sum()) + 3 * c
        print(f"   con{c}: cuda key={o[0]:.0f} hand={o[1]:.0f} link={o[2]:.0f} f={o[3:6]} b={o[6:9]} R={o[9]:.4e} A00={o[10]:.4e}")
        print(f"          orcl key={cons[c]['key']} geom1={cons[c]['geom1']} f={od.efc('efc_force')[adr:adr+3]} b={od.efc('efc_b')[adr:adr+3]} R={od.efc('efc_R')[adr]:.4e} A00={od.efc('efc_AR')[adr, adr]:.4e}")
    worst = np.argmax(np.abs(dbg[D['QA']:D['QA'] + nv] - od.qacc))
    print("   worst qacc dof", worst, dbg[D['QA'] + worst], od.qacc[worst])
    batch.debug_enable(False)


def main():
    print("device", torch.cuda.get_device_name(0))
    # ---------------- self-actuated piano
    fm, model, batch = helpers.make_native("self_actuated_piano", nenv=2, lookahead=0)
    om = O.OracleModel(fm); od = O.OracleData(om)
    rng = np.random.default_rng(0)
    nv, nu = fm.meta["nv"], fm.meta["nu"]
    q = np.zeros(nv); v = np.zeros(nv); w = np.zeros(nv)
    ctrl = rng.uniform(0, 1, nu) * fm.actuator_ctrlrange[:, 1]
    compare_substep("keys rest + random ctrl", fm, batch, od, q, v, w, ctrl)
    q = rng.uniform(-0.01, 0.1, nv); v = rng.normal(size=nv); w = rng.normal(size=nv)
    compare_substep("keys random state", fm, batch, od, q, v, w, ctrl)
    # trajectory parity, 300 substeps
    od.reset(); batch.reset()
    batch.set_state(np.zeros((2, nv)), np.zeros((2, nv)), np.zeros((2, nv)))
    od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
    maxe = 0
    for t in range(300):
        if t % 10 == 0:
            ctrl = rng.uniform(0, 1, nu) * fm.actuator_ctrlrange[:, 1]
            ct = torch.tensor(np.tile(ctrl, (2, 1)), dtype=torch.float64, device="cuda")
        batch.physics_step(ct.data_ptr(), 1)
        od.mj_step(ctrl)
        q2, v2, _, _ = batch.get_state()
        maxe = max(maxe, err(q2[0], od.qpos), err(v2[0], od.qvel) * 1e-2)
    print("self-actuated 300-substep trajectory max err", maxe)
    batch.close()

    # ---------------- hands
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2)
    om = O.OracleModel(fm); od = O.OracleData(om)
    om.set_option(maxcon=24)
    nv, nu = fm.meta["nv"], fm.meta["nu"]
    acts = helpers.twinkle_actions()
    q = np.zeros(nv); v = np.zeros(nv); w = np.zeros(nv)
    compare_substep("hands rest ctrl0", fm, batch, od, q, v, w, np.zeros(nu))
    lo, hi = fm.jnt_range[:, 0], fm.jnt_range[:, 1]
    q = lo + rng.random(nv) * (hi - lo); q[:88] = rng.uniform(0, 0.05, 88)
    v = rng.normal(size=nv) * 0.5; w = rng.normal(size=nv)
    compare_substep("hands random state", fm, batch, od, q, v, w, acts[0, :44])
    # states with contacts: roll the oracle along the fixture and compare single substeps
    od.reset()
    snaps = {}
    for t in range(100):
        od.control_step(acts[t, :44])
        if t in (5, 30, 52, 70, 92):
            snaps[t] = (od.qpos.copy(), od.qvel.copy(), od.qacc_warmstart.copy())
    for t, (q, v, w) in snaps.items():
        compare_substep(f"fixture t={t}", fm, batch, od, q, v, w, acts[t + 1, :44])
    # trajectory parity: 1000 substeps from rest following the fixture
    od.reset()
    od.qpos[:] = 0; od.qvel[:] = 0; od.qacc_warmstart[:] = 0
    batch.set_state(np.zeros((2, nv)), np.zeros((2, nv)), np.zeros((2, nv)))
    first_bad = None
    for t in range(1000):
        ctrl = acts[t // 10, :44]
        if t % 10 == 0:
            ct = torch.tensor(np.tile(ctrl, (2, 1)), dtype=torch.float64, device="cuda")
        batch.physics_step(ct.data_ptr(), 1)
        od.mj_step(ctrl)
        q2, v2, _, _ = batch.get_state()
        eq, ev = err(q2[0], od.qpos), err(v2[0], od.qvel)
        if t % 100 == 99 or (first_bad is None and eq > 1e-6):
            dg = batch.get_diag()
            print(f"   substep {t}: |dq|={eq:.3e} |dv|={ev:.3e} oracle ncon={od.ncon} niter={od.solver_niter} cuda ncon={dg['ncon'][0]} niter={dg['niter'][0]} flags={dg['flags'][0]}")
            if eq > 1e-6 and first_bad is None:
                first_bad = t
    print("hands trajectory first substep with |dq|>1e-6:", first_bad)
    # timing
    for N in (1024, 4096):
        fm, model, b2 = helpers.make_native("piano_with_shadow_hands", nenv=N)
        a = torch.tensor(np.tile(acts[:, None, :], (1, N, 1)), dtype=torch.float64, device="cuda")
        obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
        rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
        st = torch.empty(N, dtype=torch.uint8, device="cuda")
        b2.reset(None, obs.data_ptr())
        for t in range(5):
            b2.step(a[t].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(5, 45):
            b2.step(a[t].data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 40
        print(f"N={N}: {ms:.3f} ms/step -> {N / ms * 1e3:.0f} env-steps/s; reward mean {rew.mean().item():.4f}; diag niter mean {b2.get_diag()['niter'].mean():.1f}")
        b2.close()


if __name__ == "__main__":
    main()

"""The structure-of-arrays Gauss-Seidel solve (csrc/rp_pgs.cuh: groundwork for the split step pipeline, built
for the host here) against the oracle's dense mj_solPGS restatement (oracle.c: orc_solve_iterate).

The constraint problem is rebuilt from the oracle's own efc arrays in the layout the solver reads —
P[element * stride + env]: per hand M^-1, scalar friction / limit rows, coupled key rows, per contact the
3 x 3 block, the chain Jacobian and U = M^-1 Jc^T; keys without contacts solved in closed form up front — for
several environments at once (stride > 1), then every force and the sweep count must match the oracle."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from robopianist_b200.scene import cache

ROOT = Path(__file__).resolve().parents[1]
NK, NL = 88, 26
T_FRICTION, T_LIMIT, T_CONTACT = 0, 1, 2


@pytest.fixture(scope="module")
def hd(tmp_path_factory):
    so = tmp_path_factory.mktemp("hd") / "hd_pgs.so"
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", f"-I{ROOT / 'robopianist_b200' / 'csrc'}", "-o", str(so),
                           str(ROOT / "tests" / "csrc" / "hd_pgs.cpp")])
    lib = C.CDLL(str(so))
    lib.hd_pgs_solve.argtypes = [C.POINTER(C.c_double), C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_double]
    names = ["MINV", "FR", "FR_F", "LM", "LM_F", "VA", "CK", "CK_F", "CK_VA", "CON", "CON_STRIDE", "C_HAND", "C_CK", "C_DEP", "C_MU",
             "C_INVMK", "C_JK", "C_A", "C_R", "C_B", "C_F", "C_CHAIN", "C_JC", "C_U", "MAXD", "MAXCON", "NL"]
    off = {n: lib.hd_pgs_offset(i) for i, n in enumerate(names)}
    assert off["NL"] == NL
    return lib, off, lib.hd_pgs_size()


def build_problem(fm, od, off, size):
    """One environment's problem (flat array of `size` doubles) from the oracle's state after orc_solve_init."""
    nv = fm.meta["nv"]
    typ, eid = od.efc_type(), od.efc_id()
    J, R, b, AR, f0 = od.efc("efc_J"), od.efc("efc_R"), od.efc("efc_b"), od.efc("efc_AR"), od.efc("efc_force")
    loss = od.efc("efc_frictionloss")
    M = od.M.copy()
    cons = od.contacts()
    parent = fm.arrays["dof_parentid"]
    P = np.zeros(size)
    Minv = [np.linalg.inv(M[NK + NL * h: NK + NL * (h + 1), NK + NL * h: NK + NL * (h + 1)]) for h in range(2)]
    for h in range(2):
        P[off["MINV"] + h * NL * NL: off["MINV"] + (h + 1) * NL * NL] = Minv[h].ravel()
    # keys that carry contacts, in contact order (ascending key)
    crow = [i for i in range(len(typ)) if typ[i] == T_CONTACT][::3]
    ckeys = []
    for i in crow:
        k = cons[eid[i]]["key"]
        if k >= 0 and k not in ckeys:
            ckeys.append(k)
    rowmap = {}  # efc row -> ("fr", hd) / ("lm", hd) / ("ck", slot) / ("key", f_final) / ("con", c, r)
    imp0 = 0.0
    key_limit_row = {}
    for i in range(len(typ)):
        if typ[i] == T_FRICTION:
            hdof = eid[i] - NK
            P[off["FR"] + 5 * hdof: off["FR"] + 5 * hdof + 5] = [R[i], b[i], AR[i, i], 1.0 / AR[i, i], loss[i]]
            P[off["FR_F"] + hdof] = f0[i]
            rowmap[i] = ("fr", hdof)
        elif typ[i] == T_LIMIT and eid[i] >= NK:
            hdof = eid[i] - NK
            P[off["LM"] + 5 * hdof: off["LM"] + 5 * hdof + 5] = [J[i, eid[i]], R[i], b[i], AR[i, i], 1.0 / AR[i, i]]
            P[off["LM_F"] + hdof] = f0[i]
            rowmap[i] = ("lm", hdof)
        elif typ[i] == T_LIMIT:
            k = eid[i]
            key_limit_row[k] = i
            if k in ckeys:
                rowmap[i] = ("ck", ckeys.index(k))
            else:  # 1 x 1 island: one exact update, as the producer of the problem (the setup stage) does
                res = b[i] + AR[i, i] * f0[i]
                fn = max(0.0, f0[i] - res / AR[i, i])
                dl = fn - f0[i]
                imp0 -= 0.5 * dl * dl * AR[i, i] + dl * res
                rowmap[i] = ("key", fn)
    for s, k in enumerate(ckeys):
        i = key_limit_row.get(k)
        base = off["CK"] + 8 * s
        if i is None:
            P[base: base + 8] = [k, 0, 0, 0, 0, 1, 0, 1.0 / M[k, k]]
        else:
            P[base: base + 8] = [k, 1, J[i, k], R[i], b[i], AR[i, i], 1.0 / AR[i, i], 1.0 / M[k, k]]
            P[off["CK_F"] + s] = f0[i]
        P[off["CK_VA"] + s] = (J[:, k] @ f0) / M[k, k]
    for c, i in enumerate(crow):
        con = cons[eid[i]]
        hb = [np.abs(J[i:i + 3, NK + NL * h: NK + NL * (h + 1)]).sum() for h in range(2)]
        h = int(np.argmax(hb))
        assert hb[1 - h] == 0
        Jh = J[i:i + 3, NK + NL * h: NK + NL * (h + 1)]
        leaf = int(np.flatnonzero(np.abs(Jh).sum(axis=0))[-1])
        chain = []
        d = NK + NL * h + leaf
        while d >= NK + NL * h:
            chain.append(d - NK - NL * h)
            d = parent[d]
        chain = chain[::-1]
        assert len(chain) <= off["MAXD"] and set(np.flatnonzero(np.abs(Jh).sum(axis=0))) <= set(chain)
        q = off["CON"] + c * off["CON_STRIDE"]
        k = con["key"]
        P[q + off["C_HAND"]] = h
        P[q + off["C_CK"]] = ckeys.index(k) if k >= 0 else -1
        P[q + off["C_DEP"]] = len(chain) - 1
        P[q + off["C_MU"]] = con["friction"]
        P[q + off["C_INVMK"]] = 1.0 / M[k, k] if k >= 0 else 0.0
        for r in range(3):
            P[q + off["C_JK"] + r] = J[i + r, k] if k >= 0 else 0.0
            P[q + off["C_R"] + r] = R[i + r]
            P[q + off["C_B"] + r] = b[i + r]
            P[q + off["C_F"] + r] = f0[i + r]
            for dd, j in enumerate(chain):
                P[q + off["C_JC"] + r * off["MAXD"] + dd] = Jh[r, j]
            P[q + off["C_U"] + r * NL: q + off["C_U"] + (r + 1) * NL] = Minv[h] @ Jh[r]
            rowmap[i + r] = ("con", c, r)
        P[q + off["C_A"]: q + off["C_A"] + 9] = AR[i:i + 3, i:i + 3].ravel()
        P[q + off["C_CHAIN"]: q + off["C_CHAIN"] + len(chain)] = chain
    P[0], P[1], P[2] = len(crow), len(ckeys), imp0
    for h in range(2):
        P[off["VA"] + NL * h: off["VA"] + NL * (h + 1)] = Minv[h] @ (J[:, NK + NL * h: NK + NL * (h + 1)].T @ f0)
    return P, rowmap


def read_forces(P, off, rowmap, n):
    out = np.zeros(n)
    for i, m in rowmap.items():
        if m[0] == "fr":
            out[i] = P[off["FR_F"] + m[1]]
        elif m[0] == "lm":
            out[i] = P[off["LM_F"] + m[1]]
        elif m[0] == "ck":
            out[i] = P[off["CK_F"] + m[1]]
        elif m[0] == "key":
            out[i] = m[1]
        else:
            out[i] = P[off["CON"] + m[1] * off["CON_STRIDE"] + off["C_F"] + m[2]]
    return out


def test_soa_solver_matches_the_oracle(hd):
    lib, off, size = hd
    fm = cache.load_scene(gravity_compensation=True)
    om = O.OracleModel(fm)
    om.set_option(maxcon=24)
    acts = helpers.twinkle_actions()[:, :44]
    lo, hi = fm.actuator_ctrlrange[:, 0], fm.actuator_ctrlrange[:, 1]
    rng = np.random.default_rng(11)
    E = 3
    drivers = [O.OracleData(om) for _ in range(E)]
    probe = O.OracleData(om)
    for d in drivers:
        d.reset()
    scale = 1.0 / (fm.meta["meaninertia"] * fm.meta["nv"])
    checked = with_contacts = early = 0
    for t in range(34):
        ctrls = [np.clip(acts[t] + (rng.normal(0, 0.05, 44) * (hi - lo) / 2 if e else 0), lo, hi) for e in range(E)]
        if t >= 14 and t % 3 == 2:
            # the first substep of this control step, rebuilt for all environments in one strided buffer
            P = np.zeros((size, E))
            maps, finals, niters = [], [], []
            for e, d in enumerate(drivers):
                probe.qpos[:] = d.qpos; probe.qvel[:] = d.qvel; probe.qacc_warmstart[:] = d.qacc_warmstart
                probe.ctrl[:] = ctrls[e]
                probe.step1()
                probe.call("orc_smooth")
                probe.call("orc_solve_init")
                P[:, e], rm = build_problem(fm, probe, off, size)
                maps.append(rm)
                probe.call("orc_solve_iterate")
                finals.append(probe.efc("efc_force")); niters.append(probe.solver_niter)
                with_contacts += probe.ncon > 0
            flat = np.ascontiguousarray(P)
            for e in range(E):
                n = lib.hd_pgs_solve(flat.ctypes.data_as(C.POINTER(C.c_double)), E, e, fm.meta["iterations"], fm.meta["tolerance"], scale)
                got = read_forces(flat[:, e], off, maps[e], len(finals[e]))
                assert n == niters[e], (t, e, n, niters[e])
                assert np.abs(got - finals[e]).max() < 1e-9 * max(1.0, np.abs(finals[e]).max()), (t, e)
                checked += 1
                early += n < fm.meta["iterations"]
        for e, d in enumerate(drivers):
            d.control_step(ctrls[e])
    assert checked >= 18 and with_contacts >= 12

"""Host-side checks of csrc/rp_math.cuh (the header compiles for the host as well): the split contact update
(pgs_elliptic_pre / pgs_elliptic_post, used by the step kernel) is bit-identical to the single-function
restatement of mju PGS's elliptic block update (pgs_elliptic_update), and the absolute-force form of the
friction-loss sweep (RP_FR_ABS) produces the classic Gauss-Seidel iterates."""
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def test_split_elliptic_update_is_bit_identical(tmp_path):
    exe = tmp_path / "hd_elliptic"
    subprocess.check_call(["g++", "-O2", f"-I{ROOT / 'robopianist_b200' / 'csrc'}", "-o", str(exe),
                           str(ROOT / "tests" / "csrc" / "hd_elliptic.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "bad 0" in out.stdout


def test_absolute_force_friction_sweep_matches_classic():
    rng = np.random.default_rng(0)
    n = 26
    for trial in range(50):
        L = rng.normal(size=(n, n)) * 0.3 + np.eye(n)
        Minv = L @ L.T
        R = rng.uniform(0.01, 1.0, n)
        AR = np.diag(Minv) + R
        b = rng.normal(size=n)
        hi = rng.uniform(0.005, 0.05, n)
        f = np.clip(rng.normal(size=n) * 0.03, -hi, hi)
        va_other = rng.normal(size=n) * 0.1          # contribution of the non-friction rows
        # classic sweep (the kernel's default branch)
        f1, va1, imp1 = f.copy(), va_other + Minv @ f, 0.0
        for i in range(n):
            res = b[i] + R[i] * f1[i] + va1[i]
            fx = min(max(f1[i] - res / AR[i], -hi[i]), hi[i])
            dl = fx - f1[i]
            va1 += Minv[:, i] * dl
            imp1 -= dl * (0.5 * dl * AR[i] + res)
            f1[i] = fx
        # absolute-force form (RP_FR_ABS): Ah = -Minv, K_ij = Ah_ij / AR_j
        Ah = -Minv
        va = va_other + Minv @ f
        g = f - (b + R * f + va) / AR
        K = Ah / AR[None, :]                         # K[i, j]
        for j in range(n):
            g[j] -= sum(K[i, j] * f[i] for i in range(j))
        fx_own, g_own = f.copy(), np.zeros(n)
        for i in range(n):
            fx = min(max(g[i], -hi[i]), hi[i])
            fx_own[i], g_own[i] = fx, g[i]
            g[i + 1:] += K[i, i + 1:] * fx
        dl = fx_own - f
        res_own = (f - g_own) * AR
        imp2 = -np.sum(dl * (0.5 * dl * AR + res_own))
        va2 = va - Ah.T @ dl
        assert np.allclose(fx_own, f1, atol=1e-13)
        assert np.allclose(va2, va1, atol=1e-12)
        assert abs(imp1 - imp2) < 1e-12


def _hd_geom(tmp_path):
    import ctypes as C
    so = tmp_path / "hd_geom.so"
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", f"-I{ROOT / 'robopianist_b200' / 'csrc'}", "-o", str(so),
                           str(ROOT / "tests" / "csrc" / "hd_geom.cpp")])
    L = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    L.hd_sphere_box.restype = C.c_double; L.hd_sphere_box.argtypes = [dp, C.c_double, dp, dp, dp]
    L.hd_segment_box_closest_t.restype = C.c_double; L.hd_segment_box_closest_t.argtypes = [dp, dp, dp]
    L.hd_make_frame.argtypes = [dp]
    L.hd_impedance.restype = C.c_double; L.hd_impedance.argtypes = [dp, C.c_double, C.c_double]
    L.hd_ref_KB.argtypes = [dp, C.c_double, C.c_double, dp, dp]
    L.hd_elliptic_primal_force.argtypes = [dp, dp, C.c_double, C.c_double, dp]
    P = lambda a: a.ctypes.data_as(dp)
    return L, P


def _signed_point_box(p, half):
    q = np.abs(p) - half
    return np.linalg.norm(np.maximum(q, 0.0)) + min(q.max(), 0.0)


def test_kernel_geometry_helpers_against_brute_force(tmp_path):
    """The narrowphase building blocks the step kernel is compiled from (sphere_box, segment_box_closest_t,
    make_frame), host build, against brute force: signed distance, unit normal, mid-point convention, the
    closest point of a segment to a box by dense sampling, orthonormal right-handed contact frames."""
    L, P = _hd_geom(tmp_path)
    rng = np.random.default_rng(0)
    for _ in range(3000):
        half = rng.uniform(0.005, 0.08, 3)
        c = rng.normal(size=3) * 0.08
        r = rng.uniform(0.004, 0.012)
        n, p = np.zeros(3), np.zeros(3)
        dist = L.hd_sphere_box(P(c), r, P(half), P(n), P(p))
        assert abs(dist - (_signed_point_box(c, half) - r)) < 1e-12
        assert abs(np.linalg.norm(n) - 1) < 1e-12
        s_box, s_sph = p + 0.5 * dist * n, p - 0.5 * dist * n
        assert abs(_signed_point_box(s_box, half)) < 1e-12               # on the box surface
        assert abs(np.linalg.norm(s_sph - c) - r) < 1e-12                 # on the sphere surface
        assert np.linalg.norm((s_sph - r * n) - c) < 1e-12                # the normal points from the sphere to the box
    for _ in range(400):
        half = rng.uniform(0.005, 0.08, 3)
        a, b = rng.normal(size=3) * 0.08, rng.normal(size=3) * 0.08
        t = L.hd_segment_box_closest_t(P(a), P(b), P(half))
        ts = np.linspace(0, 1, 4001)
        d = np.array([np.linalg.norm(np.maximum(np.abs(a + u * (b - a)) - half, 0.0)) for u in ts])
        dt = np.linalg.norm(np.maximum(np.abs(a + t * (b - a)) - half, 0.0))
        assert 0.0 <= t <= 1.0 and dt <= d.min() + 1e-9
    for _ in range(500):
        f = np.zeros(9)
        v = rng.normal(size=3)
        f[:3] = v / np.linalg.norm(v)
        L.hd_make_frame(P(f))
        F = f.reshape(3, 3)
        assert np.abs(F @ F.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(F) - 1) < 1e-12


def test_kernel_constraint_parameters_against_the_documented_formulas(tmp_path):
    """impedance(), ref_KB() and the elliptic primal map of rp_math.cuh against numpy restatements of MuJoCo's
    documented formulas (solimp sigmoid d(r); K = 1 / (dmax^2 tc^2 dr^2), B = 2 / (dmax tc), refsafe;
    constraint-update force of an elliptic cone: zero / quadratic / projected zones)."""
    L, P = _hd_geom(tmp_path)
    rng = np.random.default_rng(1)

    def imp_ref(si, pos, margin):
        d0, d1 = np.clip(si[0], 1e-4, 0.9999), np.clip(si[1], 1e-4, 0.9999)
        width, mid, power = max(si[2], 0.0), np.clip(si[3], 1e-4, 0.9999), max(si[4], 1.0)
        if d0 == d1 or width <= 1e-15:
            return 0.5 * (d0 + d1)
        x = abs((pos - margin) / width)
        if x >= 1:
            return d1
        if x <= 0:
            return d0
        if power == 1:
            y = x
        elif x <= mid:
            y = x ** power / mid ** (power - 1)
        else:
            y = 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
        return d0 + y * (d1 - d0)

    for _ in range(2000):
        si = np.array([rng.uniform(0.3, 0.95), rng.uniform(0.9, 0.999), rng.choice([1e-4, 1e-3, 5e-3]), rng.uniform(0.2, 0.8), rng.choice([1.0, 2.0, 3.0])])
        pos = rng.normal() * 3e-3
        assert abs(L.hd_impedance(P(si), pos, 0.0) - imp_ref(si, pos, 0.0)) < 1e-12
    K, B = np.zeros(1), np.zeros(1)
    for _ in range(500):
        sr = np.array([rng.choice([0.005, 0.01, 0.02, 0.05]), rng.uniform(0.5, 1.5)])
        dmax, h = rng.uniform(0.9, 0.999), 0.005
        L.hd_ref_KB(P(sr), dmax, h, P(K), P(B))
        tc = max(sr[0], 2 * h)
        assert abs(K[0] - 1 / (dmax ** 2 * tc ** 2 * sr[1] ** 2)) < 1e-6 * K[0] and abs(B[0] - 2 / (dmax * tc)) < 1e-9 * B[0]
    f = np.zeros(3)
    zones = set()
    for _ in range(3000):
        jar = rng.normal(size=3) * np.array([1.0, 2.0, 2.0])
        impratio = rng.choice([1.0, 10.0, 3.0])
        Rn = rng.uniform(1e-3, 1e-1)
        D = np.array([1.0 / Rn, impratio / Rn, impratio / Rn])
        fr = rng.uniform(0.3, 1.5)
        mu = fr * np.sqrt((1 / D[1]) / (1 / D[0]))   # mj_makeImpedance: friction[0] * sqrt(R_t / R_n)
        L.hd_elliptic_primal_force(P(jar), P(D), mu, fr, P(f))
        N, T = jar[0] * mu, fr * np.hypot(jar[1], jar[2])
        if N >= mu * T:                       # jar inside the dual of the friction cone: no force
            assert (f == 0).all(); zones.add(0)
            assert jar[0] >= fr * np.hypot(jar[1], jar[2]) * (1 - 1e-12)
        elif mu * N + T <= 0:                 # the unconstrained minimiser -D jar is inside the friction cone
            assert np.allclose(f, -D * jar, rtol=1e-12); zones.add(1)
            assert np.hypot(f[1], f[2]) <= fr * f[0] * (1 + 1e-9)
        else:                                 # in between: the force lies on the surface of the friction cone
            assert f[0] > 0 and abs(np.hypot(f[1], f[2]) - fr * f[0]) < 1e-9 * max(1.0, f[0]); zones.add(2)
            # ... and is the minimiser of 0.5 (f - f*)' R (f - f*) over the cone surface, f* = -D jar (KKT: gradient
            # along the surface vanishes), which is what makes the map continuous across the zone borders
            fs = -D * jar
            t = f[1:] / np.hypot(f[1], f[2])
            g = (f - fs) / D                  # gradient R (f - f*)
            assert abs(g[0] + fr * (g[1:] @ t)) < 1e-7 * (abs(g).max() + 1e-12)
    assert zones == {0, 1, 2}


def test_assignment_solver_and_ot_reward_against_scipy(tmp_path):
    """lsa_rect / ot_fingering_reward of rp_math.cuh (host build) against scipy.optimize.linear_sum_assignment,
    which the reference calls (piano_with_shadow_hands.py:358): same pairs on continuous costs, same pairs on
    integer costs with many ties (the scan order and tie rule are restated), and the reward of
    _compute_ot_fingering_reward for 0..88 keys to press."""
    import ctypes as C
    from scipy.optimize import linear_sum_assignment
    sys_path_oracle = str(ROOT)
    import sys
    sys.path.insert(0, sys_path_oracle)
    from oracle.task_ref import tolerance
    so = tmp_path / "hd_ot.so"
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", f"-I{ROOT / 'robopianist_b200' / 'csrc'}", "-o", str(so),
                           str(ROOT / "tests" / "csrc" / "hd_ot.cpp")])
    L = C.CDLL(str(so))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.hd_ot_reward.restype = C.c_double
    rng = np.random.default_rng(0)
    for trial in range(600):
        nr = int(rng.integers(1, 11)); nc = int(rng.integers(nr, 40))
        cost = rng.uniform(0, 1, (nr, nc)) if trial % 2 == 0 else rng.integers(0, 4, (nr, nc)).astype(float)
        cost = np.ascontiguousarray(cost)
        col = np.zeros(nr, np.int32)
        L.hd_lsa(nr, nc, cost.ctypes.data_as(dp), col.ctypes.data_as(ip))
        r, c = linear_sum_assignment(cost)
        assert (r == np.arange(nr)).all() and (c == col).all(), (trial, c, col)
    for trial in range(300):
        K = int(rng.integers(0, 14)) if trial % 10 else 88
        fingers = rng.normal(size=(10, 3)) * 0.05
        keys = fingers[rng.integers(0, 10, K)] + rng.normal(size=(K, 3)) * rng.choice([0.002, 0.02, 0.2]) if K else np.zeros((0, 3))
        got = L.hd_ot_reward(np.ascontiguousarray(fingers).ctypes.data_as(dp), np.ascontiguousarray(keys.reshape(-1)).ctypes.data_as(dp) if K else None, K)
        if K == 0:
            assert got == 1.0
            continue
        dist = np.linalg.norm(keys[None, :, :] - fingers[:, None, :], axis=2)
        ri, ci = linear_sum_assignment(dist)
        ref = float(np.mean(tolerance(dist[ri, ci], (0, 0.01), 0.1)))
        assert abs(got - ref) < 1e-14, (trial, K, got, ref)


def _run_pairs(tmp_path, lines):
    exe = tmp_path / "hd_pairs"
    subprocess.check_call(["g++", "-O2", f"-I{ROOT / 'robopianist_b200' / 'csrc'}", "-o", str(exe),
                           str(ROOT / "tests" / "csrc" / "hd_pairs.cpp")])
    out = subprocess.run([str(exe)], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return out.stdout.strip().split("\n")


def test_hand_pair_narrowphase_is_geometrically_consistent(tmp_path):
    """sphere / capsule pairs (rp::sphere_capsule_pair, the kernel's hand : hand narrowphase): the reported distance is the
    brute-force distance between the two segments minus the radii, the normal is the unit vector between the closest
    points (geom 1 -> geom 2), the position lies mid-way between the surfaces; nothing beyond the margin."""
    rng = np.random.default_rng(4)
    cases, lines = [], []
    for i in range(400):
        kind = i % 3  # sphere-sphere, sphere-capsule, capsule-capsule
        r1, r2 = rng.uniform(0.005, 0.015, 2)
        h1 = 0.0 if kind < 2 else rng.uniform(0.005, 0.03)
        h2 = 0.0 if kind < 1 else rng.uniform(0.005, 0.03)
        z1, z2 = rng.normal(size=3), rng.normal(size=3)
        z1 /= np.linalg.norm(z1); z2 /= np.linalg.norm(z2)
        p1 = rng.normal(size=3) * 0.01
        p2 = p1 + rng.normal(size=3) * 0.02
        margin = 0.002 if i % 5 == 0 else 0.0
        cases.append((margin, p1, z1, r1, h1, p2, z2, r2, h2))
        lines.append("P " + " ".join(repr(float(x)) for x in np.hstack([margin, p1, z1, r1, h1, p2, z2, r2, h2])))
    outs = _run_pairs(tmp_path, lines)
    hits = 0
    ts = np.linspace(-1, 1, 2001)
    for (margin, p1, z1, r1, h1, p2, z2, r2, h2), line in zip(cases, outs):
        v = [float(x) for x in line.split()]
        a = p1[None] + ts[:, None] * h1 * z1[None]
        b = p2[None] + ts[:, None] * h2 * z2[None]
        d2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1) if h1 > 0 else ((a[:1, None, :] - b[None, :, :]) ** 2).sum(-1)
        ia, ib = np.unravel_index(np.argmin(d2), d2.shape)
        dmin = np.sqrt(d2[ia, ib])
        if dmin - r1 - r2 > margin + 1e-4:
            assert int(v[0]) == 0
            continue
        if int(v[0]) == 0:
            assert dmin - r1 - r2 > margin - 1e-4
            continue
        hits += 1
        dist, n, pos = v[1], np.array(v[2:5]), np.array(v[5:8])
        assert abs(dist - (dmin - r1 - r2)) < 2e-5, (dist, dmin - r1 - r2)
        assert abs(np.linalg.norm(n) - 1) < 1e-12
        want = (b[ib] - a[ia]) / dmin
        assert np.abs(n - want).max() < 5e-2 or dmin < 1e-4
        # mid-way between the surfaces along the normal, seen from both bodies' closest points
        assert np.abs(pos - (a[ia] + n * (r1 + 0.5 * dist))).max() < 1e-4
    assert hits > 100


def test_obb_overlap_matches_a_plain_separating_axis_test(tmp_path):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(9)
    cases, lines = [], []
    for i in range(300):
        Ra, Rb = Rotation.random(random_state=rng.integers(1 << 30)).as_matrix(), Rotation.random(random_state=rng.integers(1 << 30)).as_matrix()
        ha, hb = rng.uniform(0.01, 0.05, 3), rng.uniform(0.01, 0.05, 3)
        ca = rng.normal(size=3) * 0.01
        cb = ca + rng.normal(size=3) * 0.05
        cases.append((ca, Ra, ha, cb, Rb, hb))
        lines.append("B " + " ".join(repr(float(x)) for x in np.hstack([ca, Ra.ravel(), ha, cb, Rb.ravel(), hb])))
    outs = _run_pairs(tmp_path, lines)

    def overlap(ca, Ra, ha, cb, Rb, hb):
        axes = [Ra[:, i] for i in range(3)] + [Rb[:, j] for j in range(3)] + [np.cross(Ra[:, i], Rb[:, j]) for i in range(3) for j in range(3)]
        for ax in axes:
            n = np.linalg.norm(ax)
            if n < 1e-9:
                continue
            ax = ax / n
            if abs(ax @ (cb - ca)) > sum(ha[i] * abs(ax @ Ra[:, i]) for i in range(3)) + sum(hb[j] * abs(ax @ Rb[:, j]) for j in range(3)):
                return False
        return True

    got = [int(x) for x in outs]
    want = [int(overlap(*c)) for c in cases]
    assert got == want and 30 < sum(want) < 270

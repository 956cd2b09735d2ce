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

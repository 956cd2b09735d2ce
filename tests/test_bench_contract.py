"""bench.py contract checks that run without a GPU: the reference arm (CPU oracle port on the host cores) prints
one JSON line with the keys the driver reads, our arm refuses to run without CUDA (no CPU fallback), and the
algorithmic byte count matches SURVEY.md §8(d)'s bookkeeping."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["dtype"] == "f64"


def test_our_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "CUDA" in (out.stderr + out.stdout)


def test_algorithmic_bytes():
    sys.path.insert(0, str(ROOT))
    import bench
    # fp64: state read + write (3 x 140 x 8 each way), action 45, ctrl 44 read + write, obs 329, scalars
    assert bench.algorithmic_bytes_per_env_step(140, 44, 329, 45) == 2 * 3360 + 360 + 704 + 2632 + 17 + 8 + 8 == 10449

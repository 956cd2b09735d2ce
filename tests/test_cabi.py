"""The C-ABI library loads and exports every symbol include/roboplay.h declares (no compute)."""
import re
from pathlib import Path

import pytest

from robopianist_b200 import native

ROOT = Path(__file__).resolve().parents[1]


def test_header_symbols_exported():
    if not native.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    L = native.lib()
    header = (ROOT / "include" / "roboplay.h").read_text()
    declared = set(re.findall(r"\b(rp_[a-z_]+)\s*\(", header))
    assert declared == set(native.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.rp_version() == 1


def test_no_cpu_fallback():
    import helpers
    from robopianist_b200.scene import cache
    fm = cache.load_scene()
    goal, fing = helpers.song_tables()
    model = native.Model(fm, goal, fing, kind=0)
    assert (model.obs_dim, model.act_dim) == (329, 45)
    import ctypes
    assert not native.lib().rp_batch_create(model.ptr, 4, -1, ctypes.c_void_p(0))
    assert b"no CPU path" in native.lib().rp_last_error()


def test_product_does_not_import_oracle():
    pkg = ROOT / "robopianist_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        txt = p.read_text()
        assert "oracle" not in txt.replace("CPU oracle (oracle/)", "").replace("the oracle", "").lower() or p.name in ("compile.py",), p

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


def test_ot_fingering_wrapper_on_the_host_matches_the_reference_logic():
    """The per-environment body of rp_ot_fingering_kernel (csrc/rp_task.cuh: which goal row, fingertip stacking,
    key poses from the key angles, scratch layout) run on the host through a debug entry of the library, fed
    with the oracle's post-step positions and fingertip sites: equals the optimal-transport term the restated
    reference task computes for the same step (piano_with_shadow_hands.py:333-369)."""
    import ctypes as C
    import sys
    sys.path.insert(0, str(ROOT / "tests"))
    import numpy as np
    import helpers
    from oracle import task_ref
    from robopianist_b200.scene import cache
    fm = cache.load_scene(gravity_compensation=True)
    L = native.lib()
    L.rp_debug_ot_fingering_host.restype = C.c_double
    L.rp_debug_ot_fingering_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    dp = C.POINTER(C.c_double)
    for song, steps in (("TwinkleTwinkleRousseau", 45), ("CMajorScaleTwoHands", 25)):
        goal, fing = helpers.song_tables(song)
        model = native.Model(fm, goal, fing, kind=0, disable_fingering_reward=True, ot_fingering_reward=True)
        ref = task_ref.RefEnv(fm, goal, fing, disable_fingering_reward=True, ot_fingering_reward=True)
        ref.reset()
        acts = helpers.twinkle_actions()
        partial = 0
        for t in range(steps):
            ref.step(acts[t])
            q = np.ascontiguousarray(ref.d.qpos, np.float64)
            sx = np.ascontiguousarray(ref.d.site_xpos[:10], np.float64)
            got = L.rp_debug_ot_fingering_host(model.ptr, ref.t_idx, q.ctypes.data_as(dp), sx.ctypes.data_as(dp))
            want = ref.reward_terms["ot_fingering_reward"]
            assert abs(got - want) < 1e-12, (song, t, got, want)
            partial += 0.0 < want < 1.0
        assert partial > 5


def test_model_create_rejects_bad_task_arguments():
    """rp_model_create validates what it dereferences: null goal / fingering tables, episode length, lookahead range,
    task kind — NULL + rp_last_error instead of a crash."""
    import ctypes as C
    import helpers
    from robopianist_b200 import cabi
    from robopianist_b200.scene import cache
    fm = cache.load_scene()
    goal, fing = helpers.song_tables()
    L = native.lib()
    s, keep = cabi.pack_model(fm, None, None)

    def create(**edit):
        t, keep2 = cabi.pack_task(0, goal, fing)
        for k, v in edit.items():
            setattr(t, k, v)
        p = L.rp_model_create(C.byref(s), C.byref(t))
        if p:
            L.rp_model_destroy(p)
        return bool(p), L.rp_last_error().decode()

    assert create()[0]
    ok, msg = create(goal=C.POINTER(C.c_uint8)()); assert not ok and "null" in msg
    ok, msg = create(fingering=C.POINTER(C.c_int8)()); assert not ok and "null" in msg
    ok, msg = create(n_steps=0); assert not ok and "n_steps" in msg
    ok, msg = create(n_steps_lookahead=-1); assert not ok and "lookahead" in msg
    ok, msg = create(n_steps_lookahead=5000); assert not ok and "lookahead" in msg
    ok, msg = create(kind=7); assert not ok and "kind" in msg


def test_batch_create_fails_cleanly_without_a_device():
    """On a box without a usable CUDA device rp_batch_create(device >= 0) returns NULL with an error string (no crash)."""
    import ctypes
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    import helpers
    from robopianist_b200.scene import cache
    fm = cache.load_scene()
    goal, fing = helpers.song_tables()
    model = native.Model(fm, goal, fing, kind=0)
    assert not native.lib().rp_batch_create(model.ptr, 4, 0, ctypes.c_void_p(0))
    assert native.lib().rp_last_error()
    with pytest.raises(native.NativeError):
        native.Batch(model, 4)


@pytest.mark.gpu
def test_batch_create_releases_partial_allocations_on_failure():
    """An allocation failing half-way through rp_batch_create (injected with the library's test hook — never by
    asking for more memory than the box has) gives NULL + an error string, skips the remaining allocations and
    gives back what it had allocated; the library stays usable."""
    import ctypes
    import torch
    import helpers
    from robopianist_b200.scene import cache
    fm = cache.load_scene()
    goal, fing = helpers.song_tables()
    model = native.Model(fm, goal, fing, kind=0)
    L = native.lib()
    L.rp_debug_fail_alloc.argtypes = [ctypes.c_int]
    b = native.Batch(model, 4096); b.close()     # warm the context / allocator
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for k in (0, 3, 9):
        L.rp_debug_fail_alloc(k)
        with pytest.raises(native.NativeError, match="allocation failed"):
            native.Batch(model, 4096)
        L.rp_debug_fail_alloc(-1)
        torch.cuda.synchronize()
        free1, _ = torch.cuda.mem_get_info()
        assert abs(free0 - free1) < 64 << 20, (k, free0, free1)
    b = native.Batch(model, 8)
    b.close()

"""The pin to real MuJoCo that the build container cannot provide ("parity unpinned").

Skips unless `tests/golden/mujoco_reference.npz` exists — the dump `scripts/dump_mjmodel.py` writes on a machine that has
`mujoco` + `dm_control` + the reference installed (compiled mjModel fields and a 1000-substep mj_step trajectory of the
action replay with PGS / elliptic / 50 iterations).  With the file present:

  * the scene flattener's compile-time constants are diffed against mjModel: counts and orderings, masses / inertias,
    joint ranges, armature / damping (incl. the critically damped forearm slides, shadow_hand.py:299-301), friction
    loss, dof_invweight0 / body_invweight0, the inertia-box capsule fitted to the fingertip meshes
    (shadow_hand.py:144-152), actuator gains / ranges, contact parameters;
  * the CPU oracle (what the CUDA kernel is tested against to 1e-9) replays the same actions from the same state; the
    contact-free prefix must agree to 1e-9, and the whole 1000-substep rollout is compared at north_star's tolerance
    (1e-5 relative).  The capsule-box narrowphase here is NOT mjc_CapsuleBox, so the second
    comparison is expected to fail until that is restated; it is an `xfail(strict=False)` that reports the deviation.
"""
from pathlib import Path

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from robopianist_b200.scene import cache

REF = helpers.GOLDEN / "mujoco_reference.npz"
pytestmark = pytest.mark.skipif(not REF.exists(), reason="no MuJoCo dump (run scripts/dump_mjmodel.py where mujoco is installed)")


@pytest.fixture(scope="module")
def ref():
    return np.load(REF, allow_pickle=False)


@pytest.fixture(scope="module")
def fm():
    return cache.load_scene(gravity_compensation=True)


def test_counts_and_orderings(ref, fm):
    assert int(ref["nv"]) == fm.meta["nv"] == 140 and int(ref["nu"]) == fm.meta["nu"] == 44
    assert int(ref["nbody"]) == fm.meta["nbody"]
    ours = [n.split("/")[-1] for n in fm.meta["joint_names"]]
    theirs = [str(n).split("/")[-1] for n in ref["joint_names"]]
    assert ours == theirs
    assert [n.split("/")[-1] for n in fm.meta["actuator_names"]] == [str(n).split("/")[-1] for n in ref["actuator_names"]]


@pytest.mark.parametrize("field,rtol", [
    ("body_mass", 1e-9), ("body_inertia", 1e-7), ("body_ipos", 1e-9), ("jnt_range", 1e-9), ("jnt_stiffness", 1e-12),
    ("dof_armature", 1e-12), ("dof_damping", 1e-6), ("dof_frictionloss", 1e-12), ("dof_invweight0", 1e-6),
    ("body_invweight0", 1e-6), ("actuator_gainprm", 1e-12), ("actuator_ctrlrange", 1e-9), ("actuator_forcerange", 1e-9),
])
def test_compiled_constants_match_mjmodel(ref, fm, field, rtol):
    a, b = np.asarray(getattr(fm, field), float), np.asarray(ref[field], float)
    if field == "actuator_gainprm":
        b = b[:, 0]
    if field == "body_invweight0":
        a, b = a.reshape(len(a), -1)[:, 0], b[:, 0]
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.allclose(a, b, rtol=rtol, atol=1e-12), (field, np.abs(a - b).max())


def test_fingertip_capsule_fit_matches_the_compiler(ref, fm):
    ours = {n.split("/")[-1]: i for i, n in enumerate(fm.meta["geom_names"])}
    hits = 0
    for j, n in enumerate(ref["geom_names"]):
        key = str(n).split("/")[-1]
        if "distal" in key and key in ours and int(ref["geom_type"][j]) == 3:
            i = ours[key]
            assert np.allclose(fm.geom_size[i][:2], ref["geom_size"][j][:2], rtol=1e-6), key
            assert np.allclose(fm.geom_pos[i], ref["geom_pos"][j], atol=1e-7), key
            hits += 1
    assert hits == 10


def _oracle_rollout(fm, ref, n_steps):
    om = O.OracleModel(fm); om.set_option(maxcon=64)
    d = O.OracleData(om); d.reset()
    qs, vs, nc = [d.qpos.copy()], [d.qvel.copy()], [0]
    for t in range(n_steps):
        d.control_step(ref["traj_actions"][t][:44])
        qs.append(d.qpos.copy()); vs.append(d.qvel.copy()); nc.append(d.ncon)
    return np.array(qs), np.array(vs), np.array(nc)


def test_contact_free_prefix_matches_mj_step(ref, fm):
    ncon = ref["traj_ncon"]
    free = int(np.argmax(ncon > 0)) if (ncon > 0).any() else len(ncon) - 1
    free = max(free - 1, 1)
    q, v, _ = _oracle_rollout(fm, ref, free)
    assert np.abs(q - ref["traj_qpos"][:free + 1]).max() < 1e-9
    assert np.abs(v - ref["traj_qvel"][:free + 1]).max() < 1e-7


@pytest.mark.xfail(strict=False, reason="capsule-box narrowphase is this repository's own scheme, not mjc_CapsuleBox (DESIGN.md §2)")
def test_1000_substep_rollout_within_north_star_tolerance(ref, fm):
    n = min(100, len(ref["traj_qpos"]) - 1)
    q, v, nc = _oracle_rollout(fm, ref, n)
    rq = np.abs(q - ref["traj_qpos"][:n + 1]).max(axis=1) / np.maximum(1.0, np.abs(ref["traj_qpos"][:n + 1]).max(axis=1))
    print("max relative qpos deviation per control step:", rq.max(), "contacts ours / mujoco:", nc.sum(), ref["traj_ncon"][:n + 1].sum())
    assert rq.max() < 1e-5


# ---- optional dumps of the model variants (scripts/dump_mjmodel.py --hand-collisions --hands-together / --reduced-action-space)
@pytest.mark.parametrize("name,kw", [("mujoco_reference_hc.npz", dict(disable_hand_collisions=False)),
                                     ("mujoco_reference_ras.npz", dict(reduced_action_space=True))])
def test_variant_rollouts_against_mujoco(name, kw):
    """Hand : hand contacts go through restatements of MuJoCo's own sphere / capsule routines and body-pair filters, and the
    reduced action space only changes the model: with a dump of the corresponding MuJoCo rollout the oracle has to reproduce the
    contact counts exactly and the state to 1e-9 as long as no capsule : box (piano) contact is active, and
    the compile-time constants of the variant model have to agree."""
    path = helpers.GOLDEN / name
    if not path.exists():
        pytest.skip(f"no {name} (scripts/dump_mjmodel.py on a machine with MuJoCo)")
    ref = np.load(path, allow_pickle=False)
    fm = cache.load_scene(gravity_compensation=bool(ref["flags"][3]), **kw)
    assert int(ref["nv"]) == fm.meta["nv"] and int(ref["nu"]) == fm.meta["nu"]
    assert np.allclose(ref["dof_invweight0"], fm.dof_invweight0, rtol=1e-9) and np.allclose(ref["body_mass"], fm.body_mass)
    om = O.OracleModel(fm); om.set_option(maxcon=60)
    od = O.OracleData(om); od.reset()
    acts = ref["traj_actions"]
    for t in range(acts.shape[0]):
        od.control_step(acts[t, :fm.meta["nu"]])
        piano = any(fm.geom_hand[c["geom2"]] < 0 or fm.geom_hand[c["geom1"]] < 0 for c in od.contacts())
        if piano:
            break
        assert od.ncon == int(ref["traj_ncon"][t + 1]), t
        assert np.abs(od.qpos - ref["traj_qpos"][t + 1]).max() < 1e-9, t
    assert t > 5

"""Flattened scene vs the reference's structural pins: piano_test.py:35-65 (88 keys sorted),
shadow_hand_test.py:81-99 (24+2 joints, 20+2 actuators), piano_with_shadow_hands_test.py:96-117
(action layout [rh, lh, sustain]), SURVEY.md Appendix A (model facts)."""
import numpy as np

from robopianist_b200.scene import cache, compile as C


def test_model_counts_and_order():
    fm = cache.load_scene()
    m = fm.meta
    assert (m["nv"], m["nu"], m["nkey"], m["nhand"], m["ntendon"], m["nsite"]) == (140, 44, 88, 2, 8, 10)
    assert m["nsubstep"] == 10 and m["timestep"] == 0.005 and m["impratio"] == 10.0
    names = m["joint_names"]
    assert names[0].endswith("white_joint_0") and names[1].endswith("black_joint_1") and names[87].endswith("white_joint_87")
    assert [n.split("/")[-1] for n in names[88:92]] == ["forearm_tx", "forearm_ty", "rh_WRJ2", "rh_WRJ1"]
    assert names[114].startswith("lh_shadow_hand/")
    acts = [n.split("/")[-1] for n in m["actuator_names"]]
    assert acts[:2] == ["rh_A_WRJ2", "rh_A_WRJ1"] and acts[20:22] == ["forearm_tx", "forearm_ty"] and acts[22] == "lh_A_WRJ2"
    # joints_pos observable order: XML joints first, forearm dofs last (shadow_hand.py:159-183)
    assert list(fm.hand_obs_dofid[0][-2:]) == [88, 89] and fm.hand_obs_dofid[0][0] == 90


def test_key_geometry():
    lay = C.key_layout()
    ys = np.array([p[0][1] for p in lay])
    assert (np.diff(ys) > 0).all()  # keys sorted left to right
    assert sum(1 for p in lay if p[4]) == 36
    fm = cache.load_scene()
    assert np.allclose(fm.key_qmax[0], np.arctan(0.01 / 0.15)) and np.allclose(fm.key_qmax[1], np.arctan(0.008 / 0.09))
    # forearm_tx range spans the piano, offset by the hand's y position (base.py:160-164)
    j = fm.meta["joint_names"].index("rh_shadow_hand/forearm_tx")
    assert np.allclose(fm.jnt_range[j], [-0.6105 - 0.15, 0.6105 - 0.15])
    # hardened key solref = (2*dt, 1) (base.py:66)
    assert np.allclose(fm.geom_solref[fm.key_geomid[0]], [0.01, 1.0])


def test_collision_filter_and_derived_constants():
    fm = cache.load_scene()
    hand = fm.geom_hand >= 0
    assert (fm.geom_contype[hand] == 1).all() and (fm.geom_conaffinity[hand] == 0).all()  # disable_hand_collisions
    assert fm.geom_collide.sum() == 36  # 15 capsules + 3 spheres per hand
    # critical damping of the forearm slides: 2*sqrt(kp*M0)
    j = fm.meta["joint_names"].index("rh_shadow_hand/forearm_tx")
    assert np.allclose(fm.dof_damping[j], 2 * np.sqrt(300 * fm.dof_M0[j]))
    assert (fm.dof_invweight0 > 0).all() and fm.meta["meaninertia"] > 0
    sa = cache.load_scene("self_actuated_piano")
    assert sa.meta["nv"] == 88 and sa.meta["nu"] == 88 and np.allclose(sa.actuator_ctrlrange[:, 1], sa.key_qmax)

"""ctypes mirror of include/roboplay.h (``rp_flat_model`` / ``rp_task``) and struct packing.

Pure data-format code: turns a ``FlatModel`` (numpy arrays) into the POD struct the C-ABI takes.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

_I = C.POINTER(C.c_int32)
_D = C.POINTER(C.c_double)

_INT_SCALARS = ["nbody", "nv", "ngeom", "nsite", "ntendon", "nwrap", "nu", "nkey", "nhand",
                "nsubstep", "iterations", "reserved0"]
_DBL_SCALARS = ["timestep", "impratio", "tolerance", "meaninertia"]
# (name, ctype) in header order
_ARRAYS: List[Tuple[str, object]] = [
    ("body_parentid", _I), ("body_rootid", _I), ("body_dofadr", _I), ("body_dofnum", _I),
    ("body_pos", _D), ("body_quat", _D), ("body_ipos", _D), ("body_iquat", _D), ("body_mass", _D),
    ("body_inertia", _D), ("body_gravcomp", _D), ("body_invweight0", _D),
    ("jnt_type", _I), ("jnt_bodyid", _I), ("jnt_limited", _I), ("dof_parentid", _I),
    ("jnt_pos", _D), ("jnt_axis", _D), ("jnt_range", _D), ("jnt_stiffness", _D), ("jnt_springref", _D),
    ("jnt_solref", _D), ("jnt_solimp", _D), ("jnt_margin", _D),
    ("dof_damping", _D), ("dof_armature", _D), ("dof_frictionloss", _D), ("dof_solref", _D),
    ("dof_solimp", _D), ("dof_invweight0", _D),
    ("geom_type", _I), ("geom_bodyid", _I), ("geom_contype", _I), ("geom_conaffinity", _I),
    ("geom_condim", _I), ("geom_priority", _I), ("geom_key", _I), ("geom_hand", _I),
    ("geom_forearm", _I), ("geom_collide", _I),
    ("geom_pos", _D), ("geom_quat", _D), ("geom_size", _D), ("geom_friction", _D), ("geom_solref", _D),
    ("geom_solimp", _D), ("geom_solmix", _D), ("geom_margin", _D), ("geom_gap", _D),
    ("site_bodyid", _I), ("site_pos", _D),
    ("tendon_adr", _I), ("tendon_num", _I), ("wrap_dofid", _I), ("wrap_coef", _D),
    ("actuator_trntype", _I), ("actuator_trnid", _I), ("actuator_ctrllimited", _I),
    ("actuator_forcelimited", _I),
    ("actuator_gear", _D), ("actuator_gainprm", _D), ("actuator_biasprm", _D),
    ("actuator_ctrlrange", _D), ("actuator_forcerange", _D),
    ("key_dofid", _I), ("key_geomid", _I), ("key_bodyid", _I), ("key_qmax", _D),
    ("hand_obs_dofid", _I),
]


class RpFlatModel(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in _INT_SCALARS]
        + [(n, C.c_double) for n in _DBL_SCALARS]
        + [("gravity", C.c_double * 3)]
        + [(n, t) for n, t in _ARRAYS]
        + [("nobsdof", C.c_int32), ("reserved1", C.c_int32)]
        + [("exclude_bodies", _I), ("nexclude", C.c_int32), ("reserved2", C.c_int32)]
    )


class RpTask(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_steps", C.c_int32), ("n_steps_lookahead", C.c_int32),
        ("wrong_press_termination", C.c_int32), ("disable_fingering_reward", C.c_int32),
        ("disable_forearm_reward", C.c_int32), ("reward_type", C.c_int32), ("ot_fingering_reward", C.c_int32),
        ("energy_penalty_coef", C.c_double),
        ("goal", C.POINTER(C.c_uint8)), ("fingering", C.POINTER(C.c_int8)),
    ]


def pack_model(fm, iterations=None, tolerance=None):
    """FlatModel -> (RpFlatModel, keepalive list)."""
    s = RpFlatModel()
    keep = []
    meta = fm.meta
    s.nbody, s.nv, s.ngeom, s.nsite = meta["nbody"], meta["nv"], meta["ngeom"], meta["nsite"]
    s.ntendon, s.nu, s.nkey, s.nhand = meta["ntendon"], meta["nu"], meta["nkey"], meta["nhand"]
    s.nwrap = int(fm.arrays["wrap_dofid"].shape[0])
    s.nsubstep = meta["nsubstep"]
    s.iterations = int(iterations if iterations is not None else meta["iterations"])
    s.timestep, s.impratio = meta["timestep"], meta["impratio"]
    s.tolerance = float(tolerance if tolerance is not None else meta["tolerance"])
    s.meaninertia = meta["meaninertia"]
    for k in range(3):
        s.gravity[k] = float(fm.arrays["opt_gravity"][k])
    for name, typ in _ARRAYS:
        dtype = np.int32 if typ is _I else np.float64
        a = np.ascontiguousarray(fm.arrays[name], dtype=dtype).ravel()
        if a.size == 0:
            a = np.zeros(1, dtype)
        keep.append(a)
        setattr(s, name, a.ctypes.data_as(typ))
    hod = fm.arrays["hand_obs_dofid"]
    s.nobsdof = int(hod.shape[1]) if hod.ndim == 2 and hod.shape[0] else 0
    ex = np.ascontiguousarray(fm.arrays.get("exclude_bodies", np.zeros((0, 2), np.int32)), dtype=np.int32).reshape(-1, 2)
    s.nexclude = int(ex.shape[0])
    exr = ex.ravel() if ex.size else np.zeros(2, np.int32)
    keep.append(exr)
    s.exclude_bodies = exr.ctypes.data_as(_I)
    return s, keep


def pack_task(kind: int, goal: np.ndarray, fingering: np.ndarray, n_steps_lookahead: int = 1,
              wrong_press_termination: bool = False, disable_fingering_reward: bool = False,
              disable_forearm_reward: bool = False, energy_penalty_coef: float = 5e-3,
              reward_type: int = 0, ot_fingering_reward: bool = False):
    t = RpTask()
    goal = np.ascontiguousarray(goal, np.uint8)
    fingering = np.ascontiguousarray(fingering, np.int8)
    assert goal.ndim == 2 and goal.shape[1] == 89 and fingering.shape == (goal.shape[0], 88)
    t.kind, t.n_steps, t.n_steps_lookahead = int(kind), int(goal.shape[0]), int(n_steps_lookahead)
    t.wrong_press_termination = int(wrong_press_termination)
    t.disable_fingering_reward = int(disable_fingering_reward)
    t.disable_forearm_reward = int(disable_forearm_reward)
    t.reward_type = int(reward_type)
    t.ot_fingering_reward = int(ot_fingering_reward)
    t.energy_penalty_coef = float(energy_penalty_coef)
    t.goal = goal.ctypes.data_as(C.POINTER(C.c_uint8))
    t.fingering = fingering.ctypes.data_as(C.POINTER(C.c_int8))
    return t, [goal, fingering]

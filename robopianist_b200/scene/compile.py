"""Scene flattener: builds the RoboPianist scene and emits flat SoA constants (``FlatModel``).

This replaces, for this one scene, what ``dm_control.mjcf`` + MuJoCo's model compiler do for the
reference (robopianist/suite/tasks/base.py:42-197, robopianist/models/piano/piano_mjcf.py:25-402,
robopianist/models/hands/shadow_hand.py:91-311, robopianist/models/arenas/stage.py:62-69).  It
runs once on the host; the device kernels (csrc/) and the CPU oracle (oracle/) both consume the
resulting arrays through the C struct declared in include/roboplay.h.

Compile-time derived quantities that MuJoCo would compute (``dof_invweight0``,
``body_invweight0``, ``stat.meaninertia``, the ``dof_M0``-based critical damping of the forearm
slides, the inertia-box capsule fitted to the fingertip meshes) are restated here in numpy and
are UNVERIFIED against a real MuJoCo compile (none is installable in this image).
"""
from __future__ import annotations

import dataclasses
import math
import os
from pathlib import Path
from typing import Dict, Optional, Sequence

import numpy as np

from . import mathutil as mu
from . import mesh as meshfit
from . import mjcf

# MuJoCo enums (mjtJoint / mjtGeom) kept so the flat arrays read like an mjModel dump.
JNT_SLIDE, JNT_HINGE = 2, 3
GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 2, 3, 5, 6, 7
TRN_JOINT, TRN_TENDON = 0, 1
MJ_MINVAL = 1e-15

# ---- piano constants (robopianist/models/piano/piano_constants.py:22-76) -----------------
NUM_KEYS = 88
NUM_WHITE = 52
WHITE_W, WHITE_L, WHITE_H = 0.0225, 0.15, 0.0225
KEY_SPACING = 0.001
BLACK_W, BLACK_L, BLACK_H = 0.01, 0.09, 0.018
PIANO_LENGTH = NUM_WHITE * WHITE_W + (NUM_WHITE - 1) * KEY_SPACING
WHITE_X, WHITE_Z = 0.0, WHITE_H / 2
BLACK_X = -WHITE_L / 2 + BLACK_L / 2
BLACK_Z = WHITE_H + 0.0125 - BLACK_H / 2
BASE_SIZE = (0.1 / 2, PIANO_LENGTH / 2, 0.04 / 2)
BASE_POS = (-WHITE_L / 2 - 0.5 * 0.1 - 0.002, 0.0, 0.04 / 2)
WHITE_QMAX = math.atan(0.01 / WHITE_L)
BLACK_QMAX = math.atan(0.008 / BLACK_L)
KEY_MASS = {False: 0.04, True: 0.02}
KEY_SPRINGREF = -1.0 * math.pi / 180
KEY_STIFFNESS, KEY_DAMPING, KEY_ARMATURE = 2.0, 0.05, 0.001

# ---- task constants (robopianist/suite/tasks/base.py:28-37) ---------------------------------
PHYSICS_TIMESTEP = 0.005
CONTROL_TIMESTEP = 0.05
HAND_POS = {"rh": (0.4, 0.15, 0.13), "lh": (0.4, -0.15, 0.13)}
HAND_QUAT = (-1.0, -1.0, 1.0, 1.0)
FOREARM_DOFS = {  # robopianist/models/hands/shadow_hand.py:41-52
    "forearm_tx": dict(type="slide", axis=(-1, 0, 0), stiffness=300.0, range=(-1.0, 1.0)),
    "forearm_ty": dict(type="slide", axis=(0, 0, 1), stiffness=300.0, range=(0.0, 0.06)),
}
FINGERTIP_BODIES = ("thdistal", "ffdistal", "mfdistal", "rfdistal", "lfdistal")
FINGERTIP_OFFSET, THUMBTIP_OFFSET = 0.026, 0.0275
# reduced action space (robopianist/models/hands/shadow_hand.py:73-79,162-182): these joints and their actuators are
# removed from the hand (their bodies stay, rigidly attached to their parents), the THJ2 range shrinks
REDUCED_ACTION_SPACE_EXCLUDED_DOFS = ("THJ5", "THJ1", "LFJ5")
REDUCED_THUMB_RANGE = (0.0, 0.698132)

DEFAULT_ASSET_DIR = Path(os.environ.get(
    "ROBOPIANIST_SHADOW_HAND_DIR", "/root/reference/third_party/mujoco_menagerie/shadow_hand"))


def key_is_black(k: int) -> bool:
    """Key 0 is A0.  (robopianist/models/piano/piano.py:170-173)"""
    return bool([0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 0, 1][k % 12])


def key_layout():
    """Per-key (pos[3], half_size[3], qmax, mass, black) sorted by key number.

    Restates the placement loops of robopianist/models/piano/piano_mjcf.py:163-400: white keys
    on a 23.5 mm pitch, the lone A#0, then twin / triplet black groups centred on the gaps.
    """
    pitch = WHITE_W + KEY_SPACING
    white_idx = [k for k in range(NUM_KEYS) if not key_is_black(k)]
    assert len(white_idx) == NUM_WHITE
    out = {}
    for i, k in enumerate(white_idx):
        y = -PIANO_LENGTH * 0.5 + WHITE_W * 0.5 + i * pitch
        out[k] = (np.array([WHITE_X, y, WHITE_Z]), np.array([WHITE_L, WHITE_W, WHITE_H]) / 2,
                  WHITE_QMAX, KEY_MASS[False], False)
    # a black key sits on the boundary after white key number w (0-based count of whites before it)
    for k in range(NUM_KEYS):
        if not key_is_black(k):
            continue
        n_white_before = sum(1 for w in white_idx if w < k)
        if k == 1:
            y = WHITE_W + 0.5 * (-PIANO_LENGTH + KEY_SPACING)
        else:
            y = -PIANO_LENGTH * 0.5 + n_white_before * pitch
        out[k] = (np.array([BLACK_X, y, BLACK_Z]), np.array([BLACK_L, BLACK_W, BLACK_H]) / 2,
                  BLACK_QMAX, KEY_MASS[True], True)
    return [out[k] for k in range(NUM_KEYS)]


@dataclasses.dataclass
class FlatModel:
    """Flat arrays describing the compiled scene.  ``arrays`` maps field name -> ndarray."""

    arrays: Dict[str, np.ndarray]
    meta: Dict[str, object]

    def __getattr__(self, name):
        arrays = self.__dict__.get("arrays", {})
        if name in arrays:
            return arrays[name]
        meta = self.__dict__.get("meta", {})
        if name in meta:
            return meta[name]
        raise AttributeError(name)

    def save(self, path):
        np.savez_compressed(path, __meta__=np.array(repr(self.meta)), **self.arrays)

    @staticmethod
    def load(path) -> "FlatModel":
        import ast

        z = np.load(path, allow_pickle=False)
        meta = ast.literal_eval(str(z["__meta__"]))
        return FlatModel({k: z[k] for k in z.files if k != "__meta__"}, meta)


class _Builder:
    def __init__(self):
        self.bodies = []  # dicts
        self.joints = []
        self.geoms = []
        self.sites = []
        self.tendons = []
        self.actuators = []
        self.add_body("world", -1, (0, 0, 0), (1, 0, 0, 0), (0, 0, 0), (1, 0, 0, 0), 0.0, (0, 0, 0))

    def add_body(self, name, parent, pos, quat, ipos, iquat, mass, inertia, gravcomp=0.0):
        self.bodies.append(dict(name=name, parent=parent, pos=np.array(pos, float),
                                quat=mu.quat_normalize(quat), ipos=np.array(ipos, float),
                                iquat=mu.quat_normalize(iquat), mass=float(mass),
                                inertia=np.array(inertia, float), gravcomp=float(gravcomp),
                                jnts=[]))
        return len(self.bodies) - 1

    def add_joint(self, body, **kw):
        kw["body"] = body
        self.joints.append(kw)
        self.bodies[body]["jnts"].append(len(self.joints) - 1)
        return len(self.joints) - 1


def _add_piano(b: _Builder, add_actuators: bool, solref):
    base = b.add_body("piano/base", 0, BASE_POS, (1, 0, 0, 0), (0, 0, 0), (1, 0, 0, 0), 0.0, (0, 0, 0))
    gdefault = dict(contype=0, conaffinity=1, condim=3, friction=np.array([1, 0.005, 0.0001]),
                    solref=np.array(solref, float), solimp=np.array([0.9, 0.95, 0.001, 0.5, 2.0]),
                    solmix=1.0, margin=0.0, gap=0.0, priority=0)
    b.geoms.append(dict(name="piano/base_geom", type=GEOM_BOX, body=base, pos=np.zeros(3),
                        quat=np.array([1.0, 0, 0, 0]), size=np.array(BASE_SIZE), key=-2, hand=-1,
                        forearm=0, **gdefault))
    key_bodies, key_geoms, key_joints = [], [], []
    for k, (pos, half, qmax, mass, black) in enumerate(key_layout()):
        # box inertia about its centre (geom mass given, body inertial inferred from the geom)
        lx, ly, lz = 2 * half
        inertia = mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
        colour = "black" if black else "white"
        body = b.add_body(f"piano/{colour}_key_{k}", 0, pos, (1, 0, 0, 0), (0, 0, 0), (1, 0, 0, 0),
                          mass, inertia)
        b.geoms.append(dict(name=f"piano/{colour}_key_geom_{k}", type=GEOM_BOX, body=body,
                            pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), size=half.copy(),
                            key=k, hand=-1, forearm=0, **gdefault))
        j = b.add_joint(body, name=f"piano/{colour}_joint_{k}", type=JNT_HINGE,
                        pos=np.array([-half[0], 0.0, 0.0]), axis=np.array([0.0, 1.0, 0.0]),
                        limited=1, range=np.array([0.0, qmax]), stiffness=KEY_STIFFNESS,
                        springref=KEY_SPRINGREF, damping=KEY_DAMPING, armature=KEY_ARMATURE,
                        frictionloss=0.0, solreflimit=np.array([0.02, 1.0]),
                        solimplimit=np.array([0.9, 0.95, 0.001, 0.5, 2.0]),
                        solreffriction=np.array([0.02, 1.0]),
                        solimpfriction=np.array([0.9, 0.95, 0.001, 0.5, 2.0]), margin=0.0)
        key_bodies.append(body)
        key_geoms.append(len(b.geoms) - 1)
        key_joints.append(j)
        if add_actuators:
            # <general gaintype=fixed gainprm=1 biastype=none ctrlrange=[0,qmax]>
            # (robopianist/models/piano/piano_mjcf.py:56-62,80-81,99-100)
            b.actuators.append(dict(name=f"piano/{colour}_actuator_{k}", trntype=TRN_JOINT, trnid=j,
                                    gear=1.0, gain=1.0, bias=np.zeros(3), ctrllimited=1,
                                    ctrlrange=np.array([0.0, qmax]), forcelimited=0,
                                    forcerange=np.zeros(2)))
    return key_bodies, key_geoms, key_joints


def _hand_total_mass(spec: mjcf.HandSpec) -> float:
    return sum(bd.mass for bd in spec.bodies)


def _add_hand(b: _Builder, side: str, asset_dir: Path, piano_half_length: float,
              primitive_fingertip_collisions: bool, gravity_compensation: bool,
              disable_hand_collisions: bool, forearm_dofs: Sequence[str], hand_index: int,
              reduced_action_space: bool = False, attachment_yaw: float = 0.0):
    prefix = side + "_"
    spec = mjcf.parse_hand(asset_dir / ("right_hand.xml" if side == "rh" else "left_hand.xml"), prefix)
    ns = f"{side}_shadow_hand/"
    root = spec.bodies[0]
    assert root.name == prefix + "forearm"
    if reduced_action_space:  # shadow_hand.py:162-179
        for dof in REDUCED_ACTION_SPACE_EXCLUDED_DOFS:
            jname, aname = prefix + dof, prefix + "A_" + dof
            owners = [bd for bd in spec.bodies if any(j.name == jname for j in bd.joints)]
            assert len(owners) == 1, jname
            owners[0].joints = [j for j in owners[0].joints if j.name != jname]
            assert any(a.name == aname for a in spec.actuators), aname
            spec.actuators = [a for a in spec.actuators if a.name != aname]
        for bd in spec.bodies:
            for j in bd.joints:
                if j.name == prefix + "THJ2":
                    j.range = np.array(REDUCED_THUMB_RANGE, float)
        for a in spec.actuators:
            if a.name == prefix + "A_THJ2":
                a.ctrlrange = np.array(REDUCED_THUMB_RANGE, float)

    # --- forearm dofs (shadow_hand.py:272-311). The joints are appended to the root body, whose
    # childclass supplies armature / frictionloss; damping is critical for the position servo:
    # 2*sqrt(kp * dof_M0) with dof_M0 taken from a compile of the hand alone (slide => total
    # mass + armature).
    class_joint = None
    for bd in spec.bodies[1:]:
        if bd.joints:
            class_joint = bd.joints[0]
            break
    total_mass = _hand_total_mass(spec)
    y0 = HAND_POS[side][1]
    new_joints = []
    for name in forearm_dofs:
        d = FOREARM_DOFS[name]
        arm = 0.0002  # <joint armature> of the hand's root default class (right_hand.xml:9)
        m0 = total_mass + arm
        rng = np.array(d["range"], float)
        if name == "forearm_tx":  # base.py:160-164,189-194
            rng = np.array([-piano_half_length - y0, piano_half_length - y0])
        new_joints.append(mjcf.JointSpec(
            name=name, type=d["type"], pos=np.zeros(3), axis=np.array(d["axis"], float), range=rng,
            damping=2.0 * math.sqrt(d["stiffness"] * m0), armature=arm, frictionloss=0.01,
            stiffness=0.0, springref=0.0, solreflimit=class_joint.solreflimit.copy(),
            solimplimit=class_joint.solimplimit.copy(),
            solreffriction=class_joint.solreffriction.copy(),
            solimpfriction=class_joint.solimpfriction.copy(), margin=0.0))
        spec.actuators.append(mjcf.ActuatorSpec(name=name, joint=name, tendon=None,
                                                kp=d["stiffness"], ctrlrange=rng.copy(),
                                                forcerange=None))
    root.joints = new_joints + root.joints
    # hand pose (base.py:173-183)
    root.pos = np.array(HAND_POS[side], float)
    # forearms turned inwards about z by attachment_yaw degrees (base.py:175-183: mju_mulQuat(rotate_by, quaternion))
    ang = math.radians((1.0 if side == "rh" else -1.0) * float(attachment_yaw))
    rotate_by = np.array([math.cos(0.5 * ang), 0.0, 0.0, math.sin(0.5 * ang)])
    root.quat = mu.quat_normalize(mu.quat_mul(rotate_by, np.array(HAND_QUAT, float))) if attachment_yaw else mu.quat_normalize(HAND_QUAT)

    body_map = {}
    joint_map = {}
    for i, bd in enumerate(spec.bodies):
        parent = 0 if bd.parent < 0 else body_map[bd.parent]
        bid = b.add_body(ns + bd.name, parent, bd.pos, bd.quat, bd.ipos, bd.iquat, bd.mass, bd.inertia,
                         gravcomp=1.0 if gravity_compensation else 0.0)
        body_map[i] = bid
        for j in bd.joints:
            jid = b.add_joint(
                bid, name=ns + j.name, type=JNT_SLIDE if j.type == "slide" else JNT_HINGE,
                pos=j.pos.copy(), axis=j.axis.copy(), limited=1 if j.range is not None else 0,
                range=j.range.copy() if j.range is not None else np.zeros(2), stiffness=j.stiffness,
                springref=j.springref, damping=j.damping, armature=j.armature,
                frictionloss=j.frictionloss, solreflimit=j.solreflimit, solimplimit=j.solimplimit,
                solreffriction=j.solreffriction, solimpfriction=j.solimpfriction, margin=j.margin)
            joint_map[j.name] = jid
        for g in bd.geoms:
            if g.contype == 0 and g.conaffinity == 0:
                continue  # visual
            gtype = {"sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE, "cylinder": GEOM_CYLINDER,
                     "box": GEOM_BOX, "mesh": GEOM_MESH}[g.type]
            pos, quat, size = g.pos.copy(), g.quat.copy(), g.size.copy()
            if gtype == GEOM_MESH and primitive_fingertip_collisions and g.mesh.endswith("distal_pst"):
                m = spec.meshes[g.mesh]
                verts = meshfit.load_obj_vertices(spec.asset_dir / m["file"], m["scale"])
                cpos, cquat, radius, half = meshfit.fit_capsule(verts)
                pos = pos + mu.rotate(quat, cpos)
                quat = mu.quat_mul(quat, cquat)
                size = np.array([radius, half, 0.0])
                gtype = GEOM_CAPSULE
            elif gtype == GEOM_MESH:
                # not collided (see geom_collide / meta["ignored_colliders"]); kept with its bounding box in the
                # geom frame (size = half extents, pos moved to the box centre) so that tests can bound its clearance
                m = spec.meshes[g.mesh]
                verts = meshfit.load_obj_vertices(spec.asset_dir / m["file"], m["scale"])
                lo, hi = verts.min(axis=0), verts.max(axis=0)
                pos = pos + mu.rotate(quat, 0.5 * (lo + hi))
                size = 0.5 * (hi - lo)
            contype, conaff = g.contype, g.conaffinity
            if disable_hand_collisions:  # piano_with_shadow_hands.py:476-489
                contype, conaff = 1, 0
            b.geoms.append(dict(name=ns + g.name, type=gtype, body=bid, pos=pos, quat=quat, size=size,
                                contype=contype, conaffinity=conaff, condim=g.condim,
                                friction=g.friction.copy(), solref=g.solref.copy(),
                                solimp=g.solimp.copy(), solmix=g.solmix, margin=g.margin, gap=g.gap,
                                priority=g.priority, key=-1, hand=hand_index,
                                forearm=1 if i == 0 else 0))
    # fingertip sites (shadow_hand.py:191-207), order thumb, ff, mf, rf, lf
    for tip in FINGERTIP_BODIES:
        bi = spec.body_index(prefix + tip)
        off = THUMBTIP_OFFSET if tip == "thdistal" else FINGERTIP_OFFSET
        b.sites.append(dict(name=ns + tip + "_site", body=body_map[bi], pos=np.array([0.0, 0.0, off]),
                            hand=hand_index))
    tendon_map = {}
    for t in spec.tendons:
        tendon_map[t.name] = len(b.tendons)
        b.tendons.append(dict(name=ns + t.name, dofs=[joint_map[j] for j in t.joints], coefs=list(t.coefs)))
    act_ids = []
    for a in spec.actuators:
        if a.joint is not None:
            trntype, trnid = TRN_JOINT, joint_map[a.joint]
        else:
            trntype, trnid = TRN_TENDON, tendon_map[a.tendon]
        b.actuators.append(dict(
            name=ns + a.name, trntype=trntype, trnid=trnid, gear=a.gear, gain=a.kp,
            bias=np.array([0.0, -a.kp, 0.0]), ctrllimited=1 if a.ctrlrange is not None else 0,
            ctrlrange=a.ctrlrange.copy() if a.ctrlrange is not None else np.zeros(2),
            forcelimited=1 if a.forcerange is not None else 0,
            forcerange=a.forcerange.copy() if a.forcerange is not None else np.zeros(2)))
        act_ids.append(len(b.actuators) - 1)
    # observation joint order = mjcf find_all order: XML joints first, appended forearm dofs last
    # (shadow_hand.py:159-183 `_parse_mjcf_elements` runs after `_add_dofs`; children appended)
    obs_joint_names = [j.name for bd in spec.bodies for j in bd.joints if j.name not in forearm_dofs]
    obs_joint_names += list(forearm_dofs)
    obs_joints = [joint_map[n] for n in obs_joint_names]
    excludes = [(body_map[spec.body_index(a)], body_map[spec.body_index(c)]) for a, c in spec.excludes]
    return dict(impratio=spec.impratio, obs_joints=obs_joints, actuators=act_ids, excludes=excludes,
                root_body=body_map[0], joint_map=joint_map)


# ------------------------------------------------------------------------------------------
# compile-time kinematics / inertia at qpos0 (numpy restatement of MuJoCo's set0 pass)
# ------------------------------------------------------------------------------------------
def _kinematics0(bodies, joints):
    nb = len(bodies)
    xpos = np.zeros((nb, 3))
    xquat = np.zeros((nb, 4))
    xquat[0] = [1, 0, 0, 0]
    for i in range(1, nb):
        p = bodies[i]["parent"]
        xpos[i] = xpos[p] + mu.rotate(xquat[p], bodies[i]["pos"])
        xquat[i] = mu.quat_normalize(mu.quat_mul(xquat[p], bodies[i]["quat"]))
    return xpos, xquat


def _mass_matrix0(bodies, joints, xpos, xquat):
    nv = len(joints)
    nb = len(bodies)
    # dof chain per body
    chain = [[] for _ in range(nb)]
    for i in range(1, nb):
        chain[i] = list(chain[bodies[i]["parent"]]) + list(bodies[i]["jnts"])
    anchor = np.zeros((nv, 3))
    axis = np.zeros((nv, 3))
    for j, jt in enumerate(joints):
        bq = xquat[jt["body"]]
        anchor[j] = xpos[jt["body"]] + mu.rotate(bq, jt["pos"])
        axis[j] = mu.rotate(bq, jt["axis"])
    M = np.zeros((nv, nv))
    jacs = []
    for i in range(nb):
        jp = np.zeros((3, nv))
        jr = np.zeros((3, nv))
        bd = bodies[i]
        com = xpos[i] + mu.rotate(xquat[i], bd["ipos"])
        for j in chain[i]:
            if joints[j]["type"] == JNT_HINGE:
                jr[:, j] = axis[j]
                jp[:, j] = np.cross(axis[j], com - anchor[j])
            else:
                jp[:, j] = axis[j]
        jacs.append((jp, jr))
        if bd["mass"] > 0:
            R = mu.quat_to_mat(mu.quat_mul(xquat[i], bd["iquat"]))
            Iw = R @ np.diag(bd["inertia"]) @ R.T
            M += bd["mass"] * jp.T @ jp + jr.T @ Iw @ jr
    for j, jt in enumerate(joints):
        M[j, j] += jt["armature"]
    return M, jacs


def compile_scene(
    task: str = "piano_with_shadow_hands",
    primitive_fingertip_collisions: bool = True,
    gravity_compensation: bool = False,
    disable_hand_collisions: bool = True,
    forearm_dofs: Sequence[str] = ("forearm_tx", "forearm_ty"),
    physics_timestep: float = PHYSICS_TIMESTEP,
    control_timestep: float = CONTROL_TIMESTEP,
    iterations: int = 50,
    tolerance: float = 1e-8,
    asset_dir: Optional[Path] = None,
    reduced_action_space: bool = False,
    attachment_yaw: float = 0.0,
) -> FlatModel:
    """Flatten ``piano_with_shadow_hands`` or ``self_actuated_piano`` into SoA constants."""
    if task not in ("piano_with_shadow_hands", "self_actuated_piano"):
        raise ValueError(f"unknown task {task}")
    hands = task == "piano_with_shadow_hands"
    if hands and not primitive_fingertip_collisions:
        raise NotImplementedError(
            "mesh fingertip colliders (convex-mesh narrowphase) are not implemented; pass "
            "primitive_fingertip_collisions=True (reference kwarg, tasks/base.py:101)")
    asset_dir = Path(asset_dir) if asset_dir is not None else DEFAULT_ASSET_DIR
    b = _Builder()
    # key solref hardened to (2*dt, 1) (base.py:60-66)
    key_bodies, key_geoms, key_joints = _add_piano(b, add_actuators=not hands,
                                                   solref=(2 * physics_timestep, 1.0))
    impratio = 1.0
    hand_info = []
    if hands:
        for hi, side in enumerate(("rh", "lh")):
            info = _add_hand(b, side, asset_dir, BASE_SIZE[1], primitive_fingertip_collisions,
                             gravity_compensation, disable_hand_collisions, forearm_dofs, hi,
                             reduced_action_space=reduced_action_space, attachment_yaw=attachment_yaw)
            hand_info.append(info)
            impratio = info["impratio"]  # <option impratio="10"> merged into the root on attach

    bodies, joints, geoms = b.bodies, b.joints, b.geoms
    nb, nv, ng = len(bodies), len(joints), len(geoms)
    xpos0, xquat0 = _kinematics0(bodies, joints)
    M0, jacs = _mass_matrix0(bodies, joints, xpos0, xquat0)
    Minv0 = np.linalg.inv(M0)
    dof_invweight0 = np.maximum(np.diag(Minv0), MJ_MINVAL)
    body_invweight0 = np.zeros((nb, 2))
    for i in range(1, nb):
        if not any(True for _ in jacs[i][0].nonzero()[0]) and not jacs[i][1].any():
            continue
        jp, jr = jacs[i]
        A = jp @ Minv0 @ jp.T
        Bm = jr @ Minv0 @ jr.T
        body_invweight0[i, 0] = max(MJ_MINVAL, np.trace(A) / 3.0)
        body_invweight0[i, 1] = max(MJ_MINVAL, np.trace(Bm) / 3.0)
    meaninertia = float(np.trace(M0) / max(1, nv))

    # dof tree parents / body root ids
    body_rootid = np.zeros(nb, np.int32)
    for i in range(1, nb):
        p = bodies[i]["parent"]
        body_rootid[i] = i if p == 0 else body_rootid[p]
    body_dofadr = np.full(nb, -1, np.int32)
    body_dofnum = np.zeros(nb, np.int32)
    for i in range(nb):
        if bodies[i]["jnts"]:
            body_dofadr[i] = bodies[i]["jnts"][0]
            body_dofnum[i] = len(bodies[i]["jnts"])
            assert bodies[i]["jnts"] == list(range(body_dofadr[i], body_dofadr[i] + body_dofnum[i]))
    dof_parent = np.full(nv, -1, np.int32)
    for j, jt in enumerate(joints):
        bd = jt["body"]
        first = bodies[bd]["jnts"][0]
        if j > first:
            dof_parent[j] = j - 1
        else:
            p = bodies[bd]["parent"]
            while p > 0 and not bodies[p]["jnts"]:
                p = bodies[p]["parent"]
            dof_parent[j] = bodies[p]["jnts"][-1] if p > 0 else -1

    def stack(key, lst, dtype=np.float64):
        return np.ascontiguousarray(np.array([x[key] for x in lst], dtype=dtype))

    A: Dict[str, np.ndarray] = {}
    A["body_parentid"] = np.array([max(bd["parent"], 0) for bd in bodies], np.int32)
    A["body_rootid"] = body_rootid
    A["body_dofadr"] = body_dofadr
    A["body_dofnum"] = body_dofnum
    A["body_pos"] = stack("pos", bodies)
    A["body_quat"] = stack("quat", bodies)
    A["body_ipos"] = stack("ipos", bodies)
    A["body_iquat"] = stack("iquat", bodies)
    A["body_mass"] = stack("mass", bodies)
    A["body_inertia"] = stack("inertia", bodies)
    A["body_gravcomp"] = stack("gravcomp", bodies)
    A["body_invweight0"] = body_invweight0
    A["jnt_type"] = stack("type", joints, np.int32)
    A["jnt_bodyid"] = stack("body", joints, np.int32)
    A["jnt_pos"] = stack("pos", joints)
    A["jnt_axis"] = stack("axis", joints)
    A["jnt_limited"] = stack("limited", joints, np.int32)
    A["jnt_range"] = stack("range", joints)
    A["jnt_stiffness"] = stack("stiffness", joints)
    A["jnt_springref"] = stack("springref", joints)
    A["jnt_solref"] = stack("solreflimit", joints)
    A["jnt_solimp"] = stack("solimplimit", joints)
    A["jnt_margin"] = stack("margin", joints)
    A["dof_parentid"] = dof_parent
    A["dof_damping"] = stack("damping", joints)
    A["dof_armature"] = stack("armature", joints)
    A["dof_frictionloss"] = stack("frictionloss", joints)
    A["dof_solref"] = stack("solreffriction", joints)
    A["dof_solimp"] = stack("solimpfriction", joints)
    A["dof_invweight0"] = dof_invweight0
    A["dof_M0"] = np.diag(M0).copy()
    A["geom_type"] = stack("type", geoms, np.int32)
    A["geom_bodyid"] = stack("body", geoms, np.int32)
    A["geom_pos"] = stack("pos", geoms)
    A["geom_quat"] = stack("quat", geoms)
    A["geom_size"] = stack("size", geoms)
    A["geom_contype"] = stack("contype", geoms, np.int32)
    A["geom_conaffinity"] = stack("conaffinity", geoms, np.int32)
    A["geom_condim"] = stack("condim", geoms, np.int32)
    A["geom_priority"] = stack("priority", geoms, np.int32)
    A["geom_friction"] = stack("friction", geoms)
    A["geom_solref"] = stack("solref", geoms)
    A["geom_solimp"] = stack("solimp", geoms)
    A["geom_solmix"] = stack("solmix", geoms)
    A["geom_margin"] = stack("margin", geoms)
    A["geom_gap"] = stack("gap", geoms)
    A["geom_key"] = stack("key", geoms, np.int32)      # key number, -2 = piano base, -1 = not piano
    A["geom_hand"] = stack("hand", geoms, np.int32)    # 0 = right, 1 = left, -1 = not a hand geom
    A["geom_forearm"] = stack("forearm", geoms, np.int32)
    # narrowphase coverage: hand capsules / spheres against piano boxes
    supported = np.isin(A["geom_type"], [GEOM_SPHERE, GEOM_CAPSULE]) & (A["geom_hand"] >= 0)
    A["geom_collide"] = supported.astype(np.int32)
    # hand colliders of the reference model that this narrowphase does NOT collide with the piano (boxes, cylinders,
    # the forearm mesh: MuJoCo's box-box / general convex routines are not restated) — surfaced, never silent
    ignored = [i for i in range(len(geoms)) if A["geom_hand"][i] >= 0 and not supported[i]
               and (A["geom_contype"][i] | A["geom_conaffinity"][i])]
    ns = len(b.sites)
    A["site_bodyid"] = stack("body", b.sites, np.int32) if ns else np.zeros(0, np.int32)
    A["site_pos"] = stack("pos", b.sites) if ns else np.zeros((0, 3))
    nt = len(b.tendons)
    adr, num, wj, wc = [], [], [], []
    for t in b.tendons:
        adr.append(len(wj))
        num.append(len(t["dofs"]))
        wj += t["dofs"]
        wc += t["coefs"]
    A["tendon_adr"] = np.array(adr, np.int32)
    A["tendon_num"] = np.array(num, np.int32)
    A["wrap_dofid"] = np.array(wj, np.int32)
    A["wrap_coef"] = np.array(wc, np.float64)
    acts = b.actuators
    nu = len(acts)
    A["actuator_trntype"] = stack("trntype", acts, np.int32)
    A["actuator_trnid"] = stack("trnid", acts, np.int32)
    A["actuator_gear"] = stack("gear", acts)
    A["actuator_gainprm"] = stack("gain", acts)
    A["actuator_biasprm"] = stack("bias", acts)
    A["actuator_ctrllimited"] = stack("ctrllimited", acts, np.int32)
    A["actuator_ctrlrange"] = stack("ctrlrange", acts)
    A["actuator_forcelimited"] = stack("forcelimited", acts, np.int32)
    A["actuator_forcerange"] = stack("forcerange", acts)
    A["key_dofid"] = np.array(key_joints, np.int32)
    A["key_geomid"] = np.array(key_geoms, np.int32)
    A["key_bodyid"] = np.array(key_bodies, np.int32)
    A["key_qmax"] = np.array([joints[j]["range"][1] for j in key_joints])
    if hands:
        A["hand_obs_dofid"] = np.array([h["obs_joints"] for h in hand_info], np.int32)
        A["hand_actuatorid"] = np.array([h["actuators"] for h in hand_info], np.int32)
        A["hand_rootbody"] = np.array([h["root_body"] for h in hand_info], np.int32)
    else:
        A["hand_obs_dofid"] = np.zeros((0, 0), np.int32)
        A["hand_actuatorid"] = np.zeros((0, 0), np.int32)
        A["hand_rootbody"] = np.zeros(0, np.int32)
    # <contact><exclude> body pairs of the hand models (right_hand.xml:299-302); only matter when hand geoms may
    # collide with each other (disable_hand_collisions=False)
    ex = [pair for h in hand_info for pair in h["excludes"]] if hands else []
    A["exclude_bodies"] = np.array(ex, np.int32).reshape(-1, 2)
    A["opt_gravity"] = np.array([0.0, 0.0, -9.81])
    A["qpos0"] = np.zeros(nv)

    n_sub = int(round(control_timestep / physics_timestep))
    meta = dict(
        task=task, nbody=nb, nv=nv, ngeom=ng, nsite=ns, ntendon=nt, nu=nu, nkey=NUM_KEYS,
        nhand=2 if hands else 0, timestep=float(physics_timestep), control_timestep=float(control_timestep),
        nsubstep=n_sub, impratio=float(impratio), iterations=int(iterations), tolerance=float(tolerance),
        meaninertia=meaninertia, cone="elliptic", solver="PGS",
        primitive_fingertip_collisions=bool(primitive_fingertip_collisions),
        gravity_compensation=bool(gravity_compensation),
        disable_hand_collisions=bool(disable_hand_collisions),
        reduced_action_space=bool(reduced_action_space), attachment_yaw=float(attachment_yaw),
        forearm_dofs=list(forearm_dofs),
        body_names=[bd["name"] for bd in bodies], joint_names=[j["name"] for j in joints],
        geom_names=[g["name"] for g in geoms], actuator_names=[a["name"] for a in acts],
        site_names=[s["name"] for s in b.sites],
        ignored_colliders=[geoms[i]["name"] for i in ignored], n_ignored_colliders=len(ignored),
    )
    return FlatModel(A, meta)

"""Minimal MJCF reader for the Shadow-Hand menagerie files.

Only the MJCF subset used by ``right_hand.xml`` / ``left_hand.xml`` is understood: nested
``<default class>`` trees, ``childclass`` inheritance, hinge/slide joints, primitive and mesh
geoms, explicit ``<inertial>``, ``<exclude>``, fixed tendons and ``<position>`` actuators.
(reference assets: third_party/mujoco_menagerie/shadow_hand/right_hand.xml:1-345; the
reference loads them with dm_control.mjcf in robopianist/models/hands/shadow_hand.py:119-120.)

The output is a plain python description (``HandSpec``) that ``compile.py`` edits the same
way the reference edits its mjcf tree and then flattens.
"""
from __future__ import annotations

import dataclasses
import xml.etree.ElementTree as ET
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

from . import mathutil as mu

# MuJoCo built-in defaults for the attributes we consume.
_BUILTIN = {
    "joint": dict(type="hinge", axis="0 0 1", pos="0 0 0", damping="0", armature="0",
                  frictionloss="0", stiffness="0", springref="0", solreflimit="0.02 1",
                  solimplimit="0.9 0.95 0.001 0.5 2", solreffriction="0.02 1",
                  solimpfriction="0.9 0.95 0.001 0.5 2", margin="0"),
    "geom": dict(type="sphere", pos="0 0 0", quat="1 0 0 0", contype="1", conaffinity="1",
                 condim="3", friction="1 0.005 0.0001", solref="0.02 1",
                 solimp="0.9 0.95 0.001 0.5 2", solmix="1", margin="0", gap="0", priority="0",
                 size="0 0 0", group="0"),
    "position": dict(kp="1", ctrllimited="auto", forcelimited="auto", gear="1"),
    "mesh": dict(scale="1 1 1"),
}


def _floats(s) -> np.ndarray:
    return np.array([float(x) for x in str(s).split()], dtype=np.float64)


def _solimp(s) -> np.ndarray:
    """solimp may be given with 3 values; trailing (midpoint, power) default to (0.5, 2)."""
    v = _floats(s)
    full = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
    full[: len(v)] = v
    return full


@dataclasses.dataclass
class JointSpec:
    name: str
    type: str  # "hinge" | "slide"
    pos: np.ndarray
    axis: np.ndarray
    range: Optional[np.ndarray]
    damping: float
    armature: float
    frictionloss: float
    stiffness: float
    springref: float
    solreflimit: np.ndarray
    solimplimit: np.ndarray
    solreffriction: np.ndarray
    solimpfriction: np.ndarray
    margin: float


@dataclasses.dataclass
class GeomSpec:
    name: str
    type: str  # sphere capsule cylinder box mesh
    pos: np.ndarray
    quat: np.ndarray
    size: np.ndarray
    contype: int
    conaffinity: int
    condim: int
    friction: np.ndarray
    solref: np.ndarray
    solimp: np.ndarray
    solmix: float
    margin: float
    gap: float
    priority: int
    mesh: Optional[str]
    dclass: str


@dataclasses.dataclass
class SiteSpec:
    name: str
    pos: np.ndarray


@dataclasses.dataclass
class BodySpec:
    name: str
    parent: int  # index into HandSpec.bodies, -1 = attached to the world / attachment frame
    pos: np.ndarray
    quat: np.ndarray
    ipos: np.ndarray
    iquat: np.ndarray
    mass: float
    inertia: np.ndarray
    childclass: Optional[str]
    joints: List[JointSpec]
    geoms: List[GeomSpec]
    sites: List[SiteSpec]
    gravcomp: float = 0.0


@dataclasses.dataclass
class TendonSpec:
    name: str
    joints: List[str]
    coefs: List[float]


@dataclasses.dataclass
class ActuatorSpec:
    name: str
    joint: Optional[str]
    tendon: Optional[str]
    kp: float
    ctrlrange: Optional[np.ndarray]
    forcerange: Optional[np.ndarray]
    gear: float = 1.0


@dataclasses.dataclass
class HandSpec:
    model: str
    prefix: str
    impratio: float
    bodies: List[BodySpec]
    excludes: List[tuple]
    tendons: List[TendonSpec]
    actuators: List[ActuatorSpec]
    meshes: Dict[str, dict]  # name -> {file, scale}
    asset_dir: Path

    def body_index(self, name: str) -> int:
        for i, b in enumerate(self.bodies):
            if b.name == name:
                return i
        raise KeyError(name)

    def find_joint(self, name: str) -> JointSpec:
        for b in self.bodies:
            for j in b.joints:
                if j.name == name:
                    return j
        raise KeyError(name)

    def all_joints(self) -> List[JointSpec]:
        """Joints in compiled (depth-first body, then in-body) order."""
        return [j for b in self.bodies for j in b.joints]


class _Defaults:
    """class name -> tag -> attribute dict, with parent-class inheritance resolved."""

    def __init__(self, root: Optional[ET.Element]):
        self.table: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}
        if root is not None:
            self._walk(root, "main", is_top=True)

    def _walk(self, elem: ET.Element, parent_cls: str, is_top: bool = False):
        cls = "main" if is_top else elem.attrib["class"]
        if cls not in self.table:
            # start from a deep copy of the parent's resolved table
            self.table[cls] = {t: dict(a) for t, a in self.table[parent_cls].items()}
        for child in elem:
            if child.tag == "default":
                continue
            self.table[cls].setdefault(child.tag, {}).update(child.attrib)
        for child in elem:
            if child.tag == "default":
                self._walk(child, cls)

    def resolve(self, tag: str, cls: Optional[str], attrib: Dict[str, str]) -> Dict[str, str]:
        out = dict(_BUILTIN.get(tag, {}))
        out.update(self.table.get(cls or "main", {}).get(tag, {}))
        out.update({k: v for k, v in attrib.items() if k != "class"})
        return out


def parse_hand(xml_path: Path, prefix: str) -> HandSpec:
    xml_path = Path(xml_path)
    root = ET.parse(str(xml_path)).getroot()
    compiler = root.find("compiler")
    meshdir = compiler.attrib.get("meshdir", "") if compiler is not None else ""
    if compiler is not None and compiler.attrib.get("angle", "degree") != "radian":
        raise ValueError("only angle=radian models are supported")
    option = root.find("option")
    impratio = float(option.attrib.get("impratio", "1")) if option is not None else 1.0
    defaults = _Defaults(root.find("default"))

    meshes: Dict[str, dict] = {}
    asset = root.find("asset")
    if asset is not None:
        for m in asset.findall("mesh"):
            a = defaults.resolve("mesh", m.attrib.get("class"), m.attrib)
            name = a.get("name") or Path(a["file"]).stem
            meshes[name] = dict(file=a["file"], scale=_floats(a["scale"]))

    bodies: List[BodySpec] = []

    def add_body(elem: ET.Element, parent: int, inherited_cls: Optional[str]):
        childclass = elem.attrib.get("childclass", inherited_cls)
        inertial = elem.find("inertial")
        if inertial is None:
            raise ValueError(f"body {elem.attrib.get('name')} lacks an explicit <inertial>")
        iquat = mu.quat_normalize(_floats(inertial.attrib.get("quat", "1 0 0 0")))
        body = BodySpec(
            name=elem.attrib["name"],
            parent=parent,
            pos=_floats(elem.attrib.get("pos", "0 0 0")),
            quat=mu.quat_normalize(_floats(elem.attrib.get("quat", "1 0 0 0"))),
            ipos=_floats(inertial.attrib.get("pos", "0 0 0")),
            iquat=iquat,
            mass=float(inertial.attrib["mass"]),
            inertia=_floats(inertial.attrib["diaginertia"]),
            childclass=childclass,
            joints=[],
            geoms=[],
            sites=[],
        )
        idx = len(bodies)
        bodies.append(body)
        ngeom = 0
        for child in elem:
            if child.tag == "joint":
                a = defaults.resolve("joint", child.attrib.get("class", childclass), child.attrib)
                rng = _floats(a["range"]) if "range" in a else None
                body.joints.append(
                    JointSpec(
                        name=a["name"], type=a["type"], pos=_floats(a["pos"]),
                        axis=_floats(a["axis"]) / np.linalg.norm(_floats(a["axis"])),
                        range=rng, damping=float(a["damping"]), armature=float(a["armature"]),
                        frictionloss=float(a["frictionloss"]), stiffness=float(a["stiffness"]),
                        springref=float(a["springref"]), solreflimit=_floats(a["solreflimit"]),
                        solimplimit=_solimp(a["solimplimit"]),
                        solreffriction=_floats(a["solreffriction"]),
                        solimpfriction=_solimp(a["solimpfriction"]), margin=float(a["margin"]),
                    )
                )
            elif child.tag == "geom":
                cls = child.attrib.get("class", childclass)
                a = defaults.resolve("geom", cls, child.attrib)
                gtype = a["type"]
                body.geoms.append(
                    GeomSpec(
                        name=a.get("name", f"{body.name}_geom{ngeom}"), type=gtype,
                        pos=_floats(a["pos"]), quat=mu.quat_normalize(_floats(a["quat"])),
                        size=np.resize(np.append(_floats(a["size"]), [0, 0, 0]), 3),
                        contype=int(a["contype"]), conaffinity=int(a["conaffinity"]),
                        condim=int(a["condim"]), friction=_floats(a["friction"]),
                        solref=_floats(a["solref"]), solimp=_solimp(a["solimp"]),
                        solmix=float(a["solmix"]), margin=float(a["margin"]), gap=float(a["gap"]),
                        priority=int(a["priority"]), mesh=a.get("mesh"), dclass=cls or "main",
                    )
                )
                ngeom += 1
            elif child.tag == "site":
                body.sites.append(SiteSpec(child.attrib["name"], _floats(child.attrib.get("pos", "0 0 0"))))
        for child in elem:
            if child.tag == "body":
                add_body(child, idx, childclass)

    for b in root.find("worldbody").findall("body"):
        add_body(b, -1, None)

    excludes = []
    contact = root.find("contact")
    if contact is not None:
        for e in contact.findall("exclude"):
            excludes.append((e.attrib["body1"], e.attrib["body2"]))

    tendons = []
    tnode = root.find("tendon")
    if tnode is not None:
        for f in tnode.findall("fixed"):
            js = f.findall("joint")
            tendons.append(TendonSpec(f.attrib["name"], [j.attrib["joint"] for j in js],
                                      [float(j.attrib.get("coef", "1")) for j in js]))

    actuators = []
    anode = root.find("actuator")
    if anode is not None:
        for p in anode:
            if p.tag != "position":
                raise ValueError(f"unsupported actuator <{p.tag}>")
            a = defaults.resolve("position", p.attrib.get("class"), p.attrib)
            actuators.append(
                ActuatorSpec(
                    name=a["name"], joint=a.get("joint"), tendon=a.get("tendon"), kp=float(a["kp"]),
                    ctrlrange=_floats(a["ctrlrange"]) if "ctrlrange" in a else None,
                    forcerange=_floats(a["forcerange"]) if "forcerange" in a else None,
                    gear=float(str(a["gear"]).split()[0]),
                )
            )

    return HandSpec(model=root.attrib.get("model", ""), prefix=prefix, impratio=impratio,
                    bodies=bodies, excludes=excludes, tendons=tendons, actuators=actuators,
                    meshes=meshes, asset_dir=xml_path.parent / meshdir)

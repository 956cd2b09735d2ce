"""ctypes wrapper around oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED (no MuJoCo available to pin it against).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from robopianist_b200 import cabi

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    src = _HERE / "oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        L = _LIB
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.POINTER(cabi.RpFlatModel)]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_set_option.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.orc_set_record_cap.argtypes = [C.c_void_p, C.c_int]
        L.orc_data_create.restype = C.c_void_p
        L.orc_data_create.argtypes = [C.c_void_p]
        L.orc_data_destroy.argtypes = [C.c_void_p]
        for fn in ("orc_kinematics", "orc_crb", "orc_velocity", "orc_collision", "orc_make_constraint",
                   "orc_project_constraint", "orc_actuation", "orc_solve", "orc_solve_init", "orc_solve_iterate",
                   "orc_smooth", "orc_step1", "orc_step2",
                   "orc_reset"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, fn).restype = None
        L.orc_control_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.orc_rollout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int,
                                  C.POINTER(C.c_double), C.c_int]
        L.orc_ptr.restype = C.POINTER(C.c_double)
        L.orc_ptr.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_iptr.restype = C.POINTER(C.c_int)
        L.orc_iptr.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_int.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_contact_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.orc_mac_count.restype = C.c_long
        L.orc_mac_count.argtypes = [C.c_void_p]
        L.orc_pool_create.restype = C.c_void_p
        L.orc_pool_create.argtypes = [C.c_void_p, C.c_int]
        L.orc_pool_destroy.argtypes = [C.c_void_p]
        L.orc_pool_rollout.restype = C.c_long
        L.orc_pool_rollout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                       C.c_int, C.c_int]
    return _LIB


class OracleModel:
    def __init__(self, flat_model, iterations=None, tolerance=None):
        self.fm = flat_model
        s, self._keep = cabi.pack_model(flat_model, iterations, tolerance)
        self.ptr = lib().orc_model_create(C.byref(s))
        self.nv, self.nu = flat_model.meta["nv"], flat_model.meta["nu"]
        self.nbody, self.ngeom, self.nsite = (flat_model.meta[k] for k in ("nbody", "ngeom", "nsite"))
        self.nsubstep = flat_model.meta["nsubstep"]

    def set_option(self, iterations=-1, tolerance=-1.0, maxcon=-1):
        lib().orc_set_option(self.ptr, iterations, tolerance, maxcon)

    def set_record_cap(self, cap):
        """Emulate the kernel's contact-record capacity (deepest records kept; oracle.c: orc_set_record_cap)."""
        lib().orc_set_record_cap(self.ptr, int(cap))

    def __del__(self):
        try:
            lib().orc_model_destroy(self.ptr)
        except Exception:
            pass

    def rollout(self, ctrl: np.ndarray, nenv: int, nthreads: int):
        """ctrl [T, nu] broadcast to all envs or [T, nenv, nu]; returns final qpos [nenv, nv]."""
        ctrl = np.ascontiguousarray(ctrl, np.float64)
        per_env = 1 if ctrl.ndim == 3 else 0
        out = np.zeros((nenv, self.nv))
        lib().orc_rollout(self.ptr, nenv, ctrl.shape[0], ctrl.ctypes.data_as(C.POINTER(C.c_double)), per_env,
                          out.ctypes.data_as(C.POINTER(C.c_double)), nthreads)
        return out


class OraclePool:
    """Persistent per-thread workspaces for timed multi-threaded rollouts (orc_pool_*): `run` continues every
    environment from `self.state` ([nenv, 3 nv + nu]: qpos, qvel, qacc_warmstart, ctrl)."""

    def __init__(self, model: OracleModel, nenv: int, nthreads: int):
        self.model, self.nenv, self.nthreads = model, int(nenv), int(nthreads)
        self.ptr = lib().orc_pool_create(model.ptr, self.nthreads)
        self.state = np.zeros((self.nenv, 3 * model.nv + model.nu))
        self._fresh = True

    def run(self, ctrl: np.ndarray, nthreads: int = 0) -> int:
        """ctrl [T, nu] (broadcast) or [T, nenv, nu]; returns the solver multiply-adds counted during the call."""
        ctrl = np.ascontiguousarray(ctrl, np.float64)
        per_env = 1 if ctrl.ndim == 3 else 0
        dp = C.POINTER(C.c_double)
        macs = lib().orc_pool_rollout(self.ptr, self.nenv, ctrl.shape[0], ctrl.ctypes.data_as(dp), per_env,
                                      self.state.ctypes.data_as(dp), int(self._fresh), int(nthreads))
        self._fresh = False
        return int(macs)

    @property
    def qpos(self):
        return self.state[:, :self.model.nv]

    def close(self):
        if self.ptr:
            lib().orc_pool_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_SHAPES = {
    "qpos": ("nv",), "qvel": ("nv",), "qacc": ("nv",), "qacc_warmstart": ("nv",), "ctrl": ("nu",),
    "qfrc_applied": ("nv",), "xpos": ("nbody", 3), "xquat": ("nbody", 4), "xmat": ("nbody", 9),
    "xipos": ("nbody", 3), "ximat": ("nbody", 9), "xanchor": ("nv", 3), "xaxis": ("nv", 3),
    "geom_xpos": ("ngeom", 3), "geom_xmat": ("ngeom", 9), "site_xpos": ("nsite", 3), "S": ("nv", 6),
    "M": ("nv", "nv"), "qfrc_bias": ("nv",), "qfrc_passive": ("nv",), "qfrc_actuator": ("nv",),
    "qfrc_smooth": ("nv",), "qacc_smooth": ("nv",), "qfrc_constraint": ("nv",),
    "actuator_length": ("nu",), "actuator_velocity": ("nu",), "actuator_force": ("nu",),
}
_EFC = ["efc_pos", "efc_R", "efc_D", "efc_aref", "efc_b", "efc_force", "efc_vel", "efc_imp", "efc_K",
        "efc_Bd", "efc_diagApprox", "efc_frictionloss"]


class OracleData:
    """One environment's mjData-like state; numpy views alias the C memory."""

    def __init__(self, model: OracleModel):
        self.model = model
        self.ptr = lib().orc_data_create(model.ptr)
        for name, shape in _SHAPES.items():
            dims = tuple(getattr(model, s) if isinstance(s, str) else s for s in shape)
            n = int(np.prod(dims))
            p = lib().orc_ptr(self.ptr, name.encode())
            arr = np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].reshape(dims)
            setattr(self, name, arr)

    def __del__(self):
        try:
            lib().orc_data_destroy(self.ptr)
        except Exception:
            pass

    def call(self, fn):
        getattr(lib(), fn)(self.model.ptr, self.ptr)

    def reset(self):
        self.call("orc_reset")

    def step1(self):
        self.call("orc_step1")

    def step2(self):
        self.call("orc_step2")

    def mj_step(self, ctrl=None):
        """Plain mj_step from the current (qpos, qvel): forward + Euler."""
        if ctrl is not None:
            self.ctrl[:] = ctrl
        self.step1()
        self.step2()

    def control_step(self, ctrl, nsub=None):
        c = np.ascontiguousarray(ctrl, np.float64)
        lib().orc_control_step(self.model.ptr, self.ptr, c.ctypes.data_as(C.POINTER(C.c_double)),
                               nsub if nsub is not None else self.model.nsubstep)

    @property
    def ncon(self):
        return lib().orc_int(self.ptr, b"ncon")

    @property
    def nefc(self):
        return lib().orc_int(self.ptr, b"nefc")

    @property
    def solver_niter(self):
        return lib().orc_int(self.ptr, b"solver_niter")

    @property
    def con_overflow(self):
        return lib().orc_int(self.ptr, b"con_overflow")

    def efc(self, name):
        n = self.nefc
        p = lib().orc_ptr(self.ptr, name.encode())
        if name == "efc_J":
            return np.ctypeslib.as_array(p, shape=(n * self.model.nv,)).reshape(n, self.model.nv).copy()
        if name == "efc_AR":
            cap = lib().orc_int(self.ptr, b"efc_cap")
            return np.ctypeslib.as_array(p, shape=(cap * cap,)).reshape(cap, cap)[:n, :n].copy()
        return np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy()

    def efc_type(self):
        n = self.nefc
        return np.ctypeslib.as_array(lib().orc_iptr(self.ptr, b"efc_type"), shape=(max(n, 1),))[:n].copy()

    def efc_id(self):
        n = self.nefc
        return np.ctypeslib.as_array(lib().orc_iptr(self.ptr, b"efc_id"), shape=(max(n, 1),))[:n].copy()

    def contacts(self):
        out = []
        buf = (C.c_double * 18)()
        for i in range(self.ncon):
            lib().orc_contact_get(self.ptr, i, buf)
            a = np.array(buf[:])
            out.append(dict(dist=a[0], pos=a[1:4], frame=a[4:13].reshape(3, 3), geom1=int(a[13]),
                            geom2=int(a[14]), key=int(a[15]), mu=a[16], friction=a[17]))
        return out

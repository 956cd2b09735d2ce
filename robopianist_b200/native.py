"""ctypes binding of libroboplay.so (the C-ABI in include/roboplay.h).

The library is built in-tree by ``robopianist_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback: a missing library or a machine without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import cabi

_LIB = None
import os
# ROBOPLAY_LIB selects an alternative build of the same library (kernel tuning variants)
LIB_PATH = Path(os.environ.get("ROBOPLAY_LIB") or Path(__file__).resolve().parent / "libroboplay.so")

EXPORTS = [
    "rp_last_error", "rp_version", "rp_model_create", "rp_model_destroy", "rp_obs_dim", "rp_action_dim",
    "rp_reward_terms", "rp_batch_create", "rp_batch_destroy", "rp_set_option", "rp_reset", "rp_step",
    "rp_step_host", "rp_physics_step", "rp_get_state", "rp_set_state", "rp_get_diag", "rp_get_aux",
    "rp_launch_count", "rp_step_staged", "rp_host_buffers", "rp_get_metrics", "rp_midi_enable", "rp_get_midi", "rp_set_block_envs", "rp_set_song",
    "rp_randomize_hand_positions", "rp_get_hand_offsets", "rp_checkpoint_bytes", "rp_checkpoint_save", "rp_checkpoint_load",
]


BLOCK_ENVS = 3  # RP_BLOCK_ENVS of include/roboplay.h (the phase-locked launch shape)


class NativeError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the step kernel)")
        L = C.CDLL(str(LIB_PATH))
        vp, ip, dp, u8 = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.rp_last_error.restype = C.c_char_p
        L.rp_model_create.restype = vp
        L.rp_model_create.argtypes = [C.POINTER(cabi.RpFlatModel), C.POINTER(cabi.RpTask)]
        L.rp_model_destroy.argtypes = [vp]
        for fn in ("rp_obs_dim", "rp_action_dim", "rp_reward_terms"):
            getattr(L, fn).argtypes = [vp]
        L.rp_batch_create.restype = vp
        L.rp_batch_create.argtypes = [vp, C.c_int, C.c_int, vp]
        L.rp_batch_destroy.argtypes = [vp]
        L.rp_set_option.argtypes = [vp, C.c_int, C.c_double]
        L.rp_reset.argtypes = [vp, vp, vp]
        L.rp_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.rp_step_host.argtypes = [vp, dp, dp, dp, dp, u8, dp]
        L.rp_physics_step.argtypes = [vp, vp, C.c_int]
        L.rp_get_state.argtypes = [vp, dp, dp, dp, ip]
        L.rp_set_state.argtypes = [vp, dp, dp, dp, ip]
        L.rp_get_diag.argtypes = [vp, ip, ip, ip]
        L.rp_get_aux.argtypes = [vp, dp, dp, u8]
        L.rp_step_staged.argtypes = [vp]
        L.rp_host_buffers.argtypes = [vp, C.POINTER(dp), C.POINTER(dp)]
        L.rp_get_metrics.argtypes = [vp, dp]
        L.rp_midi_enable.argtypes = [vp, C.c_int]
        L.rp_set_song.argtypes = [vp, C.POINTER(C.c_uint8), C.POINTER(C.c_int8), C.c_int]
        L.rp_set_block_envs.argtypes = [vp, C.c_int]
        L.rp_randomize_hand_positions.argtypes = [vp, C.c_int, C.c_uint64]
        L.rp_get_hand_offsets.argtypes = [vp, dp]
        L.rp_checkpoint_bytes.restype = C.c_size_t
        L.rp_checkpoint_bytes.argtypes = [vp]
        L.rp_checkpoint_save.argtypes = [vp, vp, C.c_size_t]
        L.rp_checkpoint_load.argtypes = [vp, vp, C.c_size_t]
        L.rp_get_midi.argtypes = [vp, C.POINTER(C.c_uint32), ip]
        L.rp_launch_count.restype = C.c_int64
        L.rp_launch_count.argtypes = [vp]
        L.rp_debug_enable.argtypes = [vp, C.c_int]
        L.rp_debug_read.argtypes = [vp, dp]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise NativeError(lib().rp_last_error().decode())


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class Model:
    """Host-side compiled model + task tables (rp_model)."""

    def __init__(self, flat_model, goal, fingering, kind=0, iterations=None, tolerance=None, **task_kw):
        L = lib()
        s, self._keep = cabi.pack_model(flat_model, iterations, tolerance)
        t, keep2 = cabi.pack_task(kind, goal, fingering, **task_kw)
        self._keep += keep2
        self.ptr = L.rp_model_create(C.byref(s), C.byref(t))
        if not self.ptr:
            raise NativeError(L.rp_last_error().decode())
        self.flat = flat_model
        self.nv, self.nu = flat_model.meta["nv"], flat_model.meta["nu"]
        self.obs_dim, self.act_dim = L.rp_obs_dim(self.ptr), L.rp_action_dim(self.ptr)
        self.n_steps = int(goal.shape[0])

    def __del__(self):
        try:
            if self.ptr:
                lib().rp_model_destroy(self.ptr)
        except Exception:
            pass


class Batch:
    """Device-resident batch of N environments (rp_batch). Pointers are raw ints / torch data_ptr()."""

    def __init__(self, model: Model, nenv: int, device: int = 0, stream: int = 0):
        L = lib()
        self.model, self.nenv, self.device = model, int(nenv), int(device)
        self.ptr = L.rp_batch_create(model.ptr, nenv, device, C.c_void_p(stream))
        if not self.ptr:
            raise NativeError(L.rp_last_error().decode())

    def close(self):
        if self.ptr:
            lib().rp_batch_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, iterations=-1, tolerance=-1.0):
        _check(lib().rp_set_option(self.ptr, iterations, tolerance))

    def set_block_envs(self, block_envs=0):
        """Environments per thread block of the step kernel: 0 auto, 1, or BLOCK_ENVS (rp_set_block_envs)."""
        _check(lib().rp_set_block_envs(self.ptr, int(block_envs)))

    def randomize_hand_positions(self, enable=True, seed=0):
        """Per-episode random shift of the hands along y (rp_randomize_hand_positions); enable=2: shifts do not
        accumulate over episodes."""
        _check(lib().rp_randomize_hand_positions(self.ptr, int(enable), int(seed) & (2**64 - 1)))

    def get_hand_offsets(self):
        out = np.zeros(self.nenv, np.float64)
        _check(lib().rp_get_hand_offsets(self.ptr, _dp(out)))
        return out

    def checkpoint(self) -> np.ndarray:
        """Everything a later step depends on as one uint8 array (rp_checkpoint_save); `restore` continues bit for bit."""
        n = lib().rp_checkpoint_bytes(self.ptr)
        buf = np.empty(n, np.uint8)
        _check(lib().rp_checkpoint_save(self.ptr, buf.ctypes.data_as(C.c_void_p), n))
        return buf

    def restore(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, np.uint8)
        _check(lib().rp_checkpoint_load(self.ptr, blob.ctypes.data_as(C.c_void_p), blob.size))

    def reset(self, mask_ptr=None, obs_ptr=None):
        _check(lib().rp_reset(self.ptr, mask_ptr, obs_ptr))

    def step(self, action_ptr, obs_ptr, reward_ptr, discount_ptr, step_type_ptr, terms_ptr=None):
        _check(lib().rp_step(self.ptr, action_ptr, obs_ptr, reward_ptr, discount_ptr, step_type_ptr, terms_ptr))

    def _staging(self):
        """numpy views of the pinned staging buffers (rp_host_buffers)."""
        if getattr(self, "_views", None) is None:
            N, m = self.nenv, self.model
            pa, po = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
            _check(lib().rp_host_buffers(self.ptr, C.byref(pa), C.byref(po)))
            od = m.obs_dim
            act = np.ctypeslib.as_array(pa, shape=(N, m.act_dim))
            out = np.ctypeslib.as_array(po, shape=(N * (od + 7) + (N + 7) // 8,))
            st = out[N * (od + 7):].view(np.uint8)[:N]
            self._views = (act, out[:N * od].reshape(N, od), out[N * od:N * od + N], out[N * od + N:N * od + 2 * N],
                           st, out[N * od + 2 * N:N * od + 7 * N].reshape(N, 5))
        return self._views

    def step_host(self, action: np.ndarray, copy: bool = True):
        """One control step with host buffers in and out (H2D + launch + D2H + sync inside the library).
        copy=False returns views of the pinned output buffer (overwritten by the next call)."""
        act, obs, rew, dis, st, terms = self._staging()
        np.copyto(act, np.asarray(action, np.float64).reshape(act.shape))
        _check(lib().rp_step_staged(self.ptr))
        if copy:
            return obs.copy(), rew.copy(), dis.copy(), st.copy(), terms.copy()
        return obs, rew, dis, st, terms

    def set_song(self, goal: np.ndarray, fingering: np.ndarray):
        """New goal / fingering tables (uint8 [T, 89], int8 [T, 88]) for every environment; reset afterwards."""
        g = np.ascontiguousarray(goal, np.uint8); f = np.ascontiguousarray(fingering, np.int8)
        assert g.ndim == 2 and g.shape[1] == 89 and f.shape == (g.shape[0], 88)
        _check(lib().rp_set_song(self.ptr, g.ctypes.data_as(C.POINTER(C.c_uint8)), f.ctypes.data_as(C.POINTER(C.c_int8)), int(g.shape[0])))

    def physics_step(self, ctrl_ptr=None, nsub=1):
        _check(lib().rp_physics_step(self.ptr, ctrl_ptr, nsub))

    def get_state(self):
        N, nv = self.nenv, self.model.nv
        q = np.empty((N, nv)); v = np.empty((N, nv)); w = np.empty((N, nv)); t = np.empty(N, np.int32)
        _check(lib().rp_get_state(self.ptr, _dp(q), _dp(v), _dp(w), t.ctypes.data_as(C.POINTER(C.c_int32))))
        return q, v, w, t

    def set_state(self, qpos=None, qvel=None, qacc_warmstart=None, t_idx=None):
        def prep(a):
            return None if a is None else np.ascontiguousarray(a, np.float64)
        q, v, w = prep(qpos), prep(qvel), prep(qacc_warmstart)
        t = None if t_idx is None else np.ascontiguousarray(t_idx, np.int32)
        _check(lib().rp_set_state(self.ptr, _dp(q), _dp(v), _dp(w),
                                  None if t is None else t.ctypes.data_as(C.POINTER(C.c_int32))))

    def get_diag(self):
        N = self.nenv
        a, b, c = (np.empty(N, np.int32) for _ in range(3))
        ip = C.POINTER(C.c_int32)
        _check(lib().rp_get_diag(self.ptr, a.ctypes.data_as(ip), b.ctypes.data_as(ip), c.ctypes.data_as(ip)))
        return dict(ncon=a, niter=b, flags=c)

    def get_aux(self):
        N = self.nenv
        f = np.empty((N, self.model.nu)); s = np.empty((N, 10, 3)); act = np.empty((N, 88), np.uint8)
        _check(lib().rp_get_aux(self.ptr, _dp(f), _dp(s), act.ctypes.data_as(C.POINTER(C.c_uint8))))
        return dict(actuator_force=f, site_xpos=s, activation=act)

    def get_metrics(self):
        """Per-env musical metrics (rp_get_metrics): dict of [N] arrays for the last completed episode,
        plus the running sums of the current one."""
        out = np.empty((self.nenv, 16))
        _check(lib().rp_get_metrics(self.ptr, _dp(out)))
        names = ("precision", "recall", "f1", "sustain_precision", "sustain_recall", "sustain_f1")
        d = {n: out[:, 8 + i].copy() for i, n in enumerate(names)}
        d["episodes"] = out[:, 7].astype(np.int64)
        d["running_sums"], d["running_steps"] = out[:, :6].copy(), out[:, 6].astype(np.int64)
        return d

    def midi_enable(self, capacity=4096):
        """Start recording the per-substep MIDI event stream (rp_midi_enable)."""
        _check(lib().rp_midi_enable(self.ptr, int(capacity)))
        self._midi_cap = int(capacity)

    def get_midi(self, timestep=0.005):
        """Events of the running episode per env, in the reference's message order: list (len N) of lists of
        (time_seconds, kind, midi_note) with kind in {"note_on", "note_off", "sustain_on", "sustain_off"}
        (midi_note is None for pedal events; velocity of note_on is 127 as in midi_module.py:67-72)."""
        cap = self._midi_cap
        ev = np.empty((self.nenv, cap), np.uint32); cnt = np.empty(self.nenv, np.int32)
        _check(lib().rp_get_midi(self.ptr, ev.ctypes.data_as(C.POINTER(C.c_uint32)), cnt.ctypes.data_as(C.POINTER(C.c_int32))))
        if (cnt > cap).any():
            raise NativeError(f"MIDI event buffer overflow (capacity {cap}, needed {int(cnt.max())})")
        kinds = ("note_on", "note_off", "sustain_on", "sustain_off")
        out = []
        for e in range(self.nenv):
            w = np.sort(ev[e, :cnt[e]])
            out.append([(float(x >> 9) * timestep, kinds[(x >> 7) & 3], int(x & 127) + 21 if ((x >> 7) & 3) < 2 else None)
                        for x in w.tolist()])
        return out

    @property
    def launches(self):
        return int(lib().rp_launch_count(self.ptr))

    # debug stage dump (substep 0 of the next launches)
    def debug_set_record_cap(self, cap):
        """Test hook: contact-record capacity below the built-in 21 (exercises the deepest-first selection)."""
        lib().rp_debug_set_record_cap.argtypes = [C.c_void_p, C.c_int]
        _check(lib().rp_debug_set_record_cap(self.ptr, int(cap)))

    def debug_enable(self, on=True):
        _check(lib().rp_debug_enable(self.ptr, int(on)))

    def debug_read(self):
        stride = lib().rp_debug_stride()
        out = np.empty((self.nenv, stride))
        _check(lib().rp_debug_read(self.ptr, _dp(out)))
        return out

"""RoboPianist suite over the B200 step kernel — mirror of ``robopianist.suite`` (suite/__init__.py).

``load()`` keeps the reference signature (suite/__init__.py:50-59) and adds keyword-only
``num_envs`` / ``device`` / ``devices`` / solver overrides.  ``num_envs=1`` returns a ``dm_env``-shaped
``PianoEnv``; ``num_envs>1`` returns a ``BatchedPianoEnv`` working on CUDA tensors; ``devices=[0, 1, ...]`` a
``MultiDevicePianoEnv`` (one process, the batch split over the listed GPUs; one process per GPU: ``ShardedPianoEnv``).

The PIG-dataset names (``REPERTOIRE_150`` / ``ETUDE_12``) need the PIG ``.proto`` files the reference
downloads separately (cli.py:231-288); they are not distributable, so those lists are empty here and
only the programmatic debug songs (music/library.py) are registered.  Any ``.mid`` file, and any serialised
``NoteSequence`` ``.proto`` file (the PIG format; music/proto.py), can be passed through ``midi_file``.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Mapping, Optional, Sequence

from ..music import library
from .env import BatchedPianoEnv, PianoEnv, StepType, TimeStep  # noqa: F401

REPERTOIRE_150: list = []
ETUDE_12: list = []
_DEBUG_BASE_NAME = "RoboPianist-debug-{}-v0"
DEBUG = [_DEBUG_BASE_NAME.format(name) for name in library.MIDI_NAME_TO_CALLABLE]
_DEBUG_DICT = dict(zip(DEBUG, library.MIDI_NAME_TO_CALLABLE))
ALL = REPERTOIRE_150 + ETUDE_12 + DEBUG
_ALL_DICT = dict(_DEBUG_DICT)


def load(
    environment_name: str,
    midi_file: Optional[Path] = None,
    seed: Optional[int] = None,
    stretch: float = 1.0,
    shift: int = 0,
    recompile_physics: bool = False,
    legacy_step: bool = True,
    task_kwargs: Optional[Mapping[str, Any]] = None,
    *,
    num_envs: int = 1,
    device: int = 0,
    devices: Optional[Sequence[int]] = None,
    iterations: Optional[int] = None,
    tolerance: Optional[float] = None,
):
    """Loads a RoboPianist environment backed by the CUDA rollout.

    ``task_kwargs`` are those of ``PianoWithShadowHands`` / ``PianoTask`` (piano_with_shadow_hands.py:50-66, base.py:96-107).
    All of them are honoured except mesh fingertip colliders: ``primitive_fingertip_collisions=True`` has to be passed (the bare
    default raises ``NotImplementedError``).  ``disable_hand_collisions=False`` (the default) collides the sphere / capsule geoms
    of the fingers and thumbs with each other; ``reduced_action_space``, ``attachment_yaw``, ``forearm_dofs`` (tx / ty),
    ``randomize_hand_positions``, ``augmentations``, ``n_seconds_lookahead``, ``initial_buffer_time``, ``wrong_press_termination``,
    ``disable_fingering_reward`` / ``disable_forearm_reward``, ``energy_penalty_coef``, ``gravity_compensation``, ``trim_silence``,
    ``physics_timestep`` / ``control_timestep`` behave as in the reference; ``disable_colorization`` and
    ``change_color_on_activation`` only concern rendering and are accepted without effect.

    ``seed`` seeds the environment's ``numpy.random.RandomState`` (the reference's ``random_state``): the draws of the
    song augmentations (``task_kwargs['augmentations']``, suite/variations.py) and the seed of the per-environment
    streams of ``randomize_hand_positions``.  The scene is flattened once, so ``recompile_physics`` only keeps its one
    observable effect: with it the random hand shifts start from the unshifted pose every episode instead of
    accumulating (the reference shifts a model it does not recompile).  ``legacy_step=False`` is rejected: the kernel implements the reference's
    default legacy ordering (mj_step2 + mj_step1; SURVEY.md §2 table, `Physics.step`).
    """
    if not legacy_step:
        raise NotImplementedError("only legacy_step=True (the reference default) is implemented")
    if midi_file is not None:
        midi = library.load(midi_file, stretch=stretch, shift=shift)
    else:
        if environment_name not in ALL:
            raise ValueError(f"Unknown environment {environment_name}. Available environments: {ALL}")
        midi = library.load(_ALL_DICT[environment_name], stretch=stretch, shift=shift)
    if devices is not None and len(devices) > 1:   # one process, several GPUs: the batch split into contiguous ranges
        from .multidevice import MultiDevicePianoEnv
        return MultiDevicePianoEnv(midi, num_envs, devices, iterations=iterations, tolerance=tolerance, seed=seed,
                                   recompile_physics=recompile_physics, **dict(task_kwargs or {}))
    if devices is not None:
        device = int(devices[0])
    batched = BatchedPianoEnv(midi, num_envs=num_envs, device=device, iterations=iterations,
                              tolerance=tolerance, seed=seed, recompile_physics=recompile_physics,
                              **dict(task_kwargs or {}))
    return PianoEnv(batched) if num_envs == 1 else batched


__all__ = ["ALL", "DEBUG", "ETUDE_12", "REPERTOIRE_150", "load", "BatchedPianoEnv", "PianoEnv", "StepType", "TimeStep"]

"""One process driving several GPUs: the `devices=[...]` form of the batched environment.

The global batch is split into contiguous ranges (`sharding.shard_range`), one `BatchedPianoEnv` per device; a step slices
the actions, launches every device's step kernel back to back (launches are asynchronous: the devices run concurrently)
and, if asked, collects the per-step outputs on the first device with peer copies on each device's own stream.  No
collective and no second process: for one process per GPU (the layout that scales the host side too) use
`ShardedPianoEnv` under torchrun.
"""
from __future__ import annotations

from typing import Sequence

from . import sharding
from .env import BatchedPianoEnv


class MultiDevicePianoEnv:
    def __init__(self, midi, num_envs: int, devices: Sequence[int], gather: bool = True, local_factory=None, **env_kwargs):
        import torch

        self.torch = torch
        self.devices = [int(d) for d in devices]
        if not self.devices or len(set(self.devices)) != len(self.devices):
            raise ValueError("devices must be a non-empty list of distinct CUDA device indices")
        world = len(self.devices)
        if num_envs % world != 0:
            raise ValueError(f"num_envs={num_envs} must be a multiple of the number of devices {world}")
        self.num_envs = int(num_envs)
        self.ranges = [sharding.shard_range(self.num_envs, r, world) for r in range(world)]
        if local_factory is not None:   # tests: stand-ins with obs / reward / discount / step_type tensors, reset(), step()
            self.envs = [local_factory(hi - lo, lo, d) for (lo, hi), d in zip(self.ranges, self.devices)]
        else:
            self.envs = [BatchedPianoEnv(midi, num_envs=hi - lo, device=d, **env_kwargs) for (lo, hi), d in zip(self.ranges, self.devices)]
        e0 = self.envs[0]
        self.model, self.obs_dim = e0.model, e0.model.obs_dim
        self.device = e0.device
        self.gather = bool(gather)
        if self.gather:
            kw = dict(dtype=e0.obs.dtype, device=e0.device)
            self.obs = torch.empty((self.num_envs, self.obs_dim), **kw)
            self.reward = torch.empty(self.num_envs, **kw)
            self.discount = torch.empty(self.num_envs, **kw)
            self.step_type = torch.empty(self.num_envs, dtype=e0.step_type.dtype, device=e0.device)

    def action_spec(self):
        return self.envs[0].action_spec()

    def observation_spec(self):
        return self.envs[0].observation_spec()

    def _collect(self):
        """Per-device outputs -> the first device (one copy per tensor on the source device's current stream; the consumer
        synchronises the devices — `synchronize()` — or uses the per-device tensors of `self.envs`)."""
        if not self.gather:
            return [(e.obs, e.reward, e.discount, e.step_type) for e in self.envs]
        t = self.torch
        for (lo, hi), e in zip(self.ranges, self.envs):
            with t.cuda.device(e.device) if e.device.type == "cuda" else _null():
                self.obs[lo:hi].copy_(e.obs, non_blocking=True)
                self.reward[lo:hi].copy_(e.reward, non_blocking=True)
                self.discount[lo:hi].copy_(e.discount, non_blocking=True)
                self.step_type[lo:hi].copy_(e.step_type, non_blocking=True)
        return self.obs, self.reward, self.discount, self.step_type

    def reset(self):
        for e in self.envs:
            e.reset()
        return self._collect()

    def step(self, actions):
        """actions: one tensor [num_envs, act_dim] (any device: each range is copied to its device) or a list with one
        [n_local, act_dim] tensor per device."""
        parts = actions if isinstance(actions, (list, tuple)) else [actions[lo:hi] for lo, hi in self.ranges]
        if len(parts) != len(self.envs):
            raise ValueError("one action block per device expected")
        for e, a in zip(self.envs, parts):
            if a.device != e.device:
                a = a.to(e.device, non_blocking=True)
            e.step(a.contiguous())
        return self._collect()

    def synchronize(self):
        for e in self.envs:
            if e.device.type == "cuda":
                self.torch.cuda.synchronize(e.device)

    def close(self):
        for e in self.envs:
            e.close()


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

"""A global batch of environments sharded over the ranks of a ``torch.distributed`` job (one process per GPU).

Environments are independent: rank r owns the contiguous range ``shard_range(n_total, r, world)`` and
steps it with its own ``BatchedPianoEnv``; there is no collective inside the physics.  The one data-path exchange is an
all-gather of the packed per-step outputs ``[n_local, obs_dim + 3]`` (observation, reward, discount, step type) for a
consumer that wants the whole batch on every rank (a learner) — float32 by default (the observation of the reference
is consumed as float32 by its learners; the local float64 tensors stay available), issued asynchronously after the
step kernel so that it overlaps the next step's kernel, double buffered.

Works without an initialised process group (world size 1: no collective, the "global" views are the local ones).
"""
from __future__ import annotations

from typing import Optional

from . import sharding
from .env import BatchedPianoEnv


class ShardedPianoEnv:
    def __init__(self, midi, num_envs_total: int, device: Optional[int] = None, gather: bool = True,
                 gather_dtype=None, local_factory=None, **env_kwargs):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.distributed else 0
        self.world = dist.get_world_size() if self.distributed else 1
        if num_envs_total % self.world != 0:
            raise ValueError(f"num_envs_total={num_envs_total} must be a multiple of the world size {self.world}")
        self.num_envs_total = int(num_envs_total)
        self.lo, self.hi = sharding.shard_range(self.num_envs_total, self.rank, self.world)
        if local_factory is not None:   # tests: any object with obs / reward / discount / step_type tensors, reset(), step()
            self.local = local_factory(self.hi - self.lo, self.lo)
        else:
            if device is None:
                device = torch.cuda.current_device()
            self.local = BatchedPianoEnv(midi, num_envs=self.hi - self.lo, device=device, **env_kwargs)
        self.num_envs = self.local.num_envs
        self.obs_dim = self.local.model.obs_dim
        self.gather = bool(gather) and self.world > 1
        self.gather_dtype = gather_dtype or torch.float32
        self._work = [None, None]
        self._k = 0
        if self.gather:
            dev, w = self.local.device, self.obs_dim + 3
            self._packed = [torch.empty((self.num_envs, w), dtype=self.gather_dtype, device=dev) for _ in range(2)]
            self._global = [torch.empty((self.num_envs_total, w), dtype=self.gather_dtype, device=dev) for _ in range(2)]
        self.gather_bytes_per_step = (self.num_envs_total * (self.obs_dim + 3) * self._elem_size()) if self.gather else 0

    def _elem_size(self):
        return self.torch.empty((), dtype=self.gather_dtype).element_size()

    # ------------------------------------------------------------------ stepping
    def _start_gather(self):
        if not self.gather:
            return
        k = self._k & 1
        if self._work[k] is not None:      # the buffer pair about to be overwritten: its gather (two steps old) must be done
            self._work[k].wait()
        e, p = self.local, self._packed[k]
        d = self.obs_dim
        p[:, :d].copy_(e.obs)
        p[:, d].copy_(e.reward); p[:, d + 1].copy_(e.discount); p[:, d + 2].copy_(e.step_type)
        self._work[k] = self.dist.all_gather_into_tensor(self._global[k], p, async_op=True)
        self._last = k
        self._k += 1

    def reset(self):
        out = self.local.reset()
        self._start_gather()
        return out

    def step(self, actions):
        """actions: the local shard [n_local, act_dim] or the global batch [n_total, act_dim] (sliced here).  Returns
        the LOCAL float64 outputs at once; the gathered global block of this step is `gathered()`."""
        if actions.shape[0] == self.num_envs_total and self.world > 1:
            actions = actions[self.lo:self.hi]
        out = self.local.step(actions)
        self._start_gather()
        return out

    def gathered(self):
        """(obs [n_total, obs_dim], reward, discount, step_type) of the most recent step / reset, in gather_dtype, valid
        on every rank; makes the current stream wait for the collective."""
        if not self.gather:
            e = self.local
            return e.obs, e.reward, e.discount, e.step_type
        k = self._last
        if self._work[k] is not None:
            self._work[k].wait()
            self._work[k] = None
        g, d = self._global[k], self.obs_dim
        return g[:, :d], g[:, d], g[:, d + 1], g[:, d + 2]

    def close(self):
        for w in self._work:
            if w is not None:
                w.wait()
        self.local.close()

"""N>1 host logic on CPU: world_size-2 gloo run of the shard split + the single output all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from robopianist_b200.suite import sharding


def test_shard_ranges_cover_batch():
    for n, w in ((4096, 8), (10, 3), (5, 8)):
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)
    with pytest.raises(ValueError):      # unequal shards cannot go through all_gather_into_tensor
        sharding.all_gather_outputs(torch.zeros((3, 4), dtype=torch.float64), 2, 7)


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_total, rank, world)
    ids = torch.arange(lo, hi, dtype=torch.float64)
    obs = ids[:, None] * torch.ones(1, 7, dtype=torch.float64)
    blk = sharding.pack_outputs(obs, ids * 2, torch.ones_like(ids), (ids % 3).to(torch.uint8))
    full = sharding.all_gather_outputs(blk, world, n_total)
    q.put((rank, full.numpy()))
    dist.destroy_process_group()


def test_all_gather_outputs_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    for r in range(2):
        full = res[r]
        assert full.shape == (8, 10)
        assert (full[:, 0] == np.arange(8)).all() and (full[:, 7] == 2 * np.arange(8)).all()
        assert (full[:, 9] == np.arange(8) % 3).all()


class _FakeLocalEnv:
    """Stands in for BatchedPianoEnv on CPU: env i of the global batch reports obs = 1000 i + step, reward = i, ..."""

    class _M:
        obs_dim, act_dim = 5, 2

    def __init__(self, n, first):
        self.num_envs, self.model, self.device = n, self._M(), torch.device("cpu")
        self.ids = torch.arange(first, first + n, dtype=torch.float64)
        self.obs = torch.zeros((n, 5), dtype=torch.float64)
        self.reward = torch.zeros(n, dtype=torch.float64); self.discount = torch.ones(n, dtype=torch.float64)
        self.step_type = torch.zeros(n, dtype=torch.uint8)
        self.t = 0

    def _fill(self):
        self.obs[:] = (1000 * self.ids + self.t)[:, None]
        self.reward[:] = self.ids * 0.5
        self.step_type[:] = 1 if self.t else 0
        return self.obs, self.reward, self.discount, self.step_type

    def reset(self):
        self.t = 0
        return self._fill()

    def step(self, actions):
        assert actions.shape == (self.num_envs, 2) and torch.equal(actions[:, 0], self.ids)   # got its own slice
        self.t += 1
        return self._fill()

    def close(self):
        pass


def _sharded_worker(rank, world, port, q):
    from robopianist_b200.suite.sharded import ShardedPianoEnv
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env = ShardedPianoEnv(None, 8, local_factory=_FakeLocalEnv)
    env.reset()
    recs = [tuple(x.clone().numpy() for x in env.gathered())]
    glob_actions = torch.stack([torch.arange(8, dtype=torch.float64), torch.zeros(8, dtype=torch.float64)], dim=1)
    for t in range(5):     # more steps than buffers: exercises the wait-before-reuse path
        env.step(glob_actions)
        recs.append(tuple(x.clone().numpy() for x in env.gathered()))
    env.close()
    q.put((rank, recs))
    dist.destroy_process_group()


def test_sharded_env_gathers_the_global_batch_world2():
    """ShardedPianoEnv over a fake CPU env, world size 2 (gloo): global actions are sliced per rank, every rank sees the
    whole batch's packed float32 outputs after every step, in env order, across the double-buffer wrap-around."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=180) for _ in procs)
    [p.join(timeout=60) for p in procs]
    for r in range(2):
        for t, (obs, rew, dis, st) in enumerate(res[r]):
            assert obs.dtype == np.float32 and obs.shape == (8, 5)
            assert (obs[:, 0] == 1000 * np.arange(8) + t).all() and (rew == 0.5 * np.arange(8)).all()
            assert (dis == 1).all() and (st == (1 if t else 0)).all()


def test_multi_device_env_splits_and_collects_on_cpu_stand_ins():
    """MultiDevicePianoEnv (`suite.load(..., devices=[...])`) over CPU stand-ins of the per-device environments: contiguous
    ranges, actions sliced per range (one tensor or one block per device), outputs collected in global order."""
    import torch
    from robopianist_b200.suite.multidevice import MultiDevicePianoEnv

    class Fake:
        def __init__(self, n, lo, dev):
            self.num_envs, self.lo = n, lo
            self.device = torch.device("cpu")
            self.model = type("M", (), dict(obs_dim=5, act_dim=3))()
            self.obs = torch.zeros((n, 5), dtype=torch.float64); self.reward = torch.zeros(n, dtype=torch.float64)
            self.discount = torch.ones(n, dtype=torch.float64); self.step_type = torch.zeros(n, dtype=torch.uint8)
            self.closed = False

        def reset(self):
            self.obs[:] = (self.lo + torch.arange(self.num_envs, dtype=torch.float64))[:, None]
            self.step_type.zero_()

        def step(self, a):
            assert a.shape == (self.num_envs, 3)
            self.obs[:, :3] = a; self.reward[:] = a.sum(1); self.step_type.fill_(1)

        def action_spec(self):
            return "spec"

        def observation_spec(self):
            return "ospec"

        def close(self):
            self.closed = True

    env = MultiDevicePianoEnv(None, 12, devices=[0, 1, 2], local_factory=lambda n, lo, d: Fake(n, lo, d))
    assert env.ranges == [(0, 4), (4, 8), (8, 12)] and env.action_spec() == "spec"
    obs, rew, dis, st = env.reset()
    assert obs[:, 0].tolist() == list(map(float, range(12))) and st.sum() == 0
    a = torch.arange(36, dtype=torch.float64).reshape(12, 3)
    obs, rew, dis, st = env.step(a)
    assert torch.equal(obs[:, :3], a) and torch.equal(rew, a.sum(1)) and (st == 1).all()
    obs2, *_ = env.step([a[0:4] + 1, a[4:8] + 1, a[8:12] + 1])
    assert torch.equal(obs2[:, :3], a + 1)
    with pytest.raises(ValueError):
        env.step([a[0:4], a[4:8]])
    with pytest.raises(ValueError):
        MultiDevicePianoEnv(None, 10, devices=[0, 1, 2], local_factory=lambda n, lo, d: Fake(n, lo, d))
    with pytest.raises(ValueError):
        MultiDevicePianoEnv(None, 12, devices=[0, 0], local_factory=lambda n, lo, d: Fake(n, lo, d))
    per_dev = MultiDevicePianoEnv(None, 12, devices=[0, 1, 2], gather=False, local_factory=lambda n, lo, d: Fake(n, lo, d))
    outs = per_dev.reset()
    assert len(outs) == 3 and outs[1][0][0, 0] == 4.0
    env.close()
    assert all(e.closed for e in env.envs)

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

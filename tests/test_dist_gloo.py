"""CPU, world_size 2 over gloo: batch sharding + the (epe_sum, count) all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bflow_b200 import dist as bdist


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    flow = torch.randn(6, 2, 8, 8, generator=g)
    tgt = torch.randn(6, 2, 8, 8, generator=g)
    valid = torch.rand(6, 8, 8, generator=g) > 0.3
    mine = [bdist.shard_batch(t, rank, world) for t in (flow, tgt, valid)]
    s, n = bdist.epe_sum_count(*mine)
    mean, cnt, table = bdist.gather_epe(s, n)
    q.put((rank, mean, cnt, table.tolist()))
    dist.destroy_process_group()


def test_two_rank_epe_gather_equals_single_process():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    flow = torch.randn(6, 2, 8, 8, generator=g)
    tgt = torch.randn(6, 2, 8, 8, generator=g)
    valid = torch.rand(6, 8, 8, generator=g) > 0.3
    s, n = bdist.epe_sum_count(flow, tgt, valid)
    for rank, mean, cnt, table in res:
        assert cnt == int(n)
        assert abs(mean - float(s) / int(n)) < 1e-12
        assert len(table) == world
    assert res[0][3] == res[1][3]

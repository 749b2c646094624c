"""world_size-2 gloo test of the data-parallel host logic (relightable_nr_b200/parallel.py): view sharding and the
single flat-bucket gradient all-reduce == mean of the per-rank gradients."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relightable_nr_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    grads = [torch.randn(7, 3, generator=g), torch.randn(11, generator=g), None, torch.randn(2, 2, 2, generator=g)]
    parallel.allreduce_mean_(grads)
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    parallel.allreduce_flat_mean_(flat)
    if rank == 0:
        torch.save({'grads': [t for t in grads if t is not None], 'flat': flat}, out)
    dist.destroy_process_group()


def test_allreduce_mean_matches_average(tmp_path):
    world = 2
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    exp = None
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        gs = [torch.randn(7, 3, generator=g), torch.randn(11, generator=g), torch.randn(2, 2, 2, generator=g)]
        exp = gs if exp is None else [a + b for a, b in zip(exp, gs)]
    for a, b in zip(got['grads'], exp):
        assert torch.allclose(a, b / world, atol=1e-6)
    assert torch.allclose(got['flat'], torch.arange(10, dtype=torch.float32) * 1.5)


def test_shard_views_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in parallel.shard_views(720, r, world))
        assert seen == list(range(720))
        sizes = [len(parallel.shard_views(720, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1

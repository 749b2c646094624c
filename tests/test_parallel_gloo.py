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


def _script_worker(rank, world, port, out):
    """A miniature 'unchanged single-process training script' run under install_script_hooks."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    parallel.install_script_hooks(seed=3, bucket_bytes=64)
    # ---- the "script": plain DataLoader(shuffle=True), plain Adam, no rank logic ----
    from torch.utils.data import DataLoader, TensorDataset
    torch.manual_seed(0)
    data = torch.arange(12, dtype=torch.float32).view(12, 1)
    loader = DataLoader(TensorDataset(data), batch_size=1, shuffle=True)
    model = torch.nn.Sequential(torch.nn.Linear(1, 4), torch.nn.Linear(4, 1))
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    seen, grads = [], None
    for (x,) in loader:
        seen.append(int(x.item()))
        opt.zero_grad()
        loss = (model(x) ** 2).sum()
        loss.backward()
        if grads is None:
            # what the hooks must produce: the mean over ranks of the local gradients (captured before step() averages them)
            local = [p.grad.clone() for p in model.parameters()]
        opt.step()
        if grads is None:
            grads = [p.grad.clone() for p in model.parameters()]
    torch.save({'seen': seen, 'local': local, 'avg': grads, 'w': [p.detach().clone() for p in model.parameters()]}, out % rank)
    dist.destroy_process_group()


def test_unchanged_script_becomes_data_parallel(tmp_path):
    world = 2
    out = str(tmp_path / 'r%d.pt')
    mp.spawn(_script_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r = [torch.load(out % i) for i in range(world)]
    # disjoint halves of one shared permutation
    assert len(r[0]['seen']) == len(r[1]['seen']) == 6
    assert sorted(r[0]['seen'] + r[1]['seen']) == list(range(12))
    # first step: averaged gradient == mean of the two local gradients, identical on both ranks
    for a0, a1, l0, l1 in zip(r[0]['avg'], r[1]['avg'], r[0]['local'], r[1]['local']):
        assert torch.allclose(a0, (l0 + l1) / 2, atol=1e-6)
        assert torch.allclose(a0, a1, atol=1e-7)
    # identical updates on every rank -> replicas stay in sync for the whole epoch
    for w0, w1 in zip(r[0]['w'], r[1]['w']):
        assert torch.allclose(w0, w1, atol=1e-6)

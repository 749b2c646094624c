"""Data-parallel plumbing for the per-view step: views are the independent units.

One process per GPU.  Inference shards views round-robin with no collective; training adds exactly one exchange per step,
a sum all-reduce of the flat gradient buffer followed by a 1/world scale (the mean over the world's views -- what a
single-process reference run over those views would average to).  BatchNorm statistics stay per rank, like the reference
(batch 1 per device).  Works with any torch.distributed backend: NCCL over NVLink on the box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_views(num_views, rank, world):
    """Indices of the views rank ``rank`` owns: i with i % world == rank (test_rnr.py:265 loop, sharded)."""
    return list(range(rank, num_views, world))


def flatten_grads(tensors):
    """One contiguous fp32 bucket holding ``tensors`` back to back, plus the (offset, numel) table to scatter it back."""
    table, o = [], 0
    for t in tensors:
        table.append((o, t.numel()))
        o += t.numel()
    flat = torch.cat([t.reshape(-1) for t in tensors]) if tensors else torch.zeros(0)
    return flat, table


def unflatten_into(flat, table, tensors):
    for (o, n), t in zip(table, tensors):
        t.copy_(flat[o:o + n].view_as(t))


def allreduce_mean_(tensors, group=None):
    """In-place mean over ranks of a list of gradient tensors, as ONE collective on one flat bucket."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = [t for t in tensors if t is not None]
    flat, table = flatten_grads(tensors)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / dist.get_world_size(group))
    unflatten_into(flat, table, tensors)


def allreduce_flat_mean_(flat, group=None, async_op=False):
    """Same on an already-flat buffer (the U-Net engine's ``grad_flat``): no packing copies."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if not async_op:
        flat.mul_(1.0 / dist.get_world_size(group))
    return work

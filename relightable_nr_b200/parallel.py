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


# ----------------------------------------------------------------------------------------------------------------------
# Data parallelism for the reference's UNCHANGED single-process training scripts: the launcher
# (relightable_nr_b200/run.py --dp) starts one process per GPU and calls install_script_hooks() before runpy.  train_rnr.py has
# no sampler, no collective and no rank logic, so everything is injected from outside:
#   * training DataLoaders (shuffle=True) get a rank-strided sampler with a shared seed -> every rank draws a different view of
#     the same permutation (1 view per GPU per step, the script's own constraint);
#   * every Parameter that receives a gradient is averaged over the ranks: post-accumulate-grad hooks enqueue an asynchronous
#     all-reduce per parameter bucket while loss.backward() (train_rnr.py:618) is still running, and a global optimizer-step
#     pre-hook waits for them and scales by 1/world before optimizerG.step() (:622);
#   * ranks > 0 get a no-op SummaryWriter and do not write checkpoints.
# ----------------------------------------------------------------------------------------------------------------------
class RankStridedSampler(torch.utils.data.Sampler):
    """Indices rank, rank + world, ... of one permutation shared by all ranks (seed + epoch); every rank gets the same count."""

    def __init__(self, data_source, rank, world, shuffle=True, seed=0):
        self.n, self.rank, self.world, self.shuffle, self.seed, self.epoch = len(data_source), rank, world, shuffle, seed, 0

    def __iter__(self):
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=g).tolist()
        else:
            order = list(range(self.n))
        self.epoch += 1
        per = -(-self.n // self.world)
        order = (order + order[:per * self.world - self.n])[:per * self.world]       # pad by wrapping, like DistributedSampler
        return iter(order[self.rank::self.world])

    def __len__(self):
        return -(-self.n // self.world)


class _GradAverager:
    """Bucketed asynchronous gradient all-reduce driven by autograd hooks."""

    def __init__(self, bucket_bytes=64 << 20, group=None):
        self.bucket_bytes, self.group = bucket_bytes, group
        self.pending = []          # [(work, flat, [grads])]
        self.cur, self.cur_bytes = [], 0
        self.hooked = set()
        self.checked = False       # bucket layout verified across ranks (first step only)

    def attach(self, params):
        for p in params:
            if p.requires_grad and id(p) not in self.hooked:
                self.hooked.add(id(p))
                p.register_post_accumulate_grad_hook(self._on_grad)

    def _on_grad(self, p):
        if p.grad is None:
            return
        self.cur.append(p.grad)
        self.cur_bytes += p.grad.numel() * p.grad.element_size()
        if self.cur_bytes >= self.bucket_bytes:
            self._flush()

    def _flush(self):
        if not self.cur:
            return
        grads, self.cur, self.cur_bytes = self.cur, [], 0
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((work, flat, grads))

    def finish(self):
        """Wait for every bucket, scale by 1/world and scatter back into the .grad tensors (called before optimizer.step())."""
        self._flush()
        world = dist.get_world_size(self.group)
        if not self.checked:
            # every rank must have produced the same buckets in the same order, or the collectives above pair up wrongly /
            # hang: compare (bucket count, total elements) across ranks once, on the first step
            sig = torch.tensor([len(self.pending), sum(f.numel() for _, f, _ in self.pending)], dtype=torch.int64,
                               device=self.pending[0][1].device if self.pending else 'cpu')
            lo, hi = sig.clone(), sig.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            if not torch.equal(lo, hi):
                raise RuntimeError('data-parallel ranks produced different gradient buckets (min %s, max %s): every rank must '
                                   'differentiate the same parameters in the same order' % (lo.tolist(), hi.tolist()))
            self.checked = True
        for work, flat, grads in self.pending:
            work.wait()
            flat.mul_(1.0 / world)
            o = 0
            for g in grads:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
        self.pending = []


def broadcast_params_(tensors, src=0, group=None):
    """In-place broadcast of ``tensors`` (parameters or buffers) from rank ``src``, one flat bucket per dtype."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    by_dtype = {}
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.numel() > 0:
            by_dtype.setdefault((t.dtype, t.device), []).append(t)
    with torch.no_grad():
        for ts in by_dtype.values():
            flat = torch.cat([t.detach().reshape(-1) for t in ts])
            dist.broadcast(flat, src=src, group=group)
            o = 0
            for t in ts:
                t.detach().copy_(flat[o:o + t.numel()].view_as(t))
                o += t.numel()


def broadcast_module_state_(modules, src=0, group=None):
    """Parameters AND buffers (BatchNorm running statistics, spectral-norm u / v, ...) of ``modules`` from rank ``src``."""
    ts = []
    for m in modules:
        ts += [p for p in m.parameters()] + [b for b in m.buffers()]
    broadcast_params_(ts, src, group)


def install_script_hooks(rank=None, world=None, seed=0, bucket_bytes=64 << 20):
    """Make an unchanged single-process training script data-parallel (see above).  Requires an initialised process group.
    Returns the gradient averager (its ``attach`` is applied automatically to the parameters of every optimizer created later)."""
    import torch.optim
    import torch.utils.data as tud
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    avg = _GradAverager(bucket_bytes)
    if getattr(tud.DataLoader, '__rnr_dp__', False):
        return avg
    orig_loader = tud.DataLoader

    class DataLoader(orig_loader):
        __rnr_dp__ = True

        def __init__(self, dataset, *a, **k):
            shuffle = k.get('shuffle', a[1] if len(a) > 1 else False)
            if shuffle and k.get('sampler') is None and k.get('batch_sampler') is None and world > 1:
                k['sampler'] = RankStridedSampler(dataset, rank, world, shuffle=True, seed=seed)
                k['shuffle'] = False
                a = a[:1] + (False,) + a[2:] if len(a) > 1 else a
            super().__init__(dataset, *a, **k)

    tud.DataLoader = DataLoader
    torch.utils.data.DataLoader = DataLoader
    # parameters are discovered when the script builds its optimizer (train_rnr.py:376): hook them there
    orig_init = torch.optim.Optimizer.__init__
    state = {'reseeded': False}

    def opt_init(self, params, defaults):
        orig_init(self, params, defaults)
        for grp in self.param_groups:
            # the scripts never seed (train_rnr.py / train_dnr.py have no torch.manual_seed): every rank built its own random
            # U-Net / GCN.  Replicas must start identical -- gradients are averaged and only rank 0 checkpoints -- so rank 0's
            # parameters are broadcast to everyone the moment the script hands them to its optimizer (train_rnr.py:376).
            broadcast_params_(grp['params'])
            avg.attach(grp['params'])
        if not state['reseeded']:
            # construction-time randomness was shared (the launcher seeds every rank alike); from here on each rank draws its
            # own dropout masks / stochastic kNN dilation
            state['reseeded'] = True
            torch.manual_seed(1000003 * (seed + 1) + rank)

    torch.optim.Optimizer.__init__ = opt_init
    from torch.optim.optimizer import register_optimizer_step_pre_hook
    register_optimizer_step_pre_hook(lambda opt, args, kwargs: avg.finish())
    if rank != 0:
        import sys
        import types
        null = types.ModuleType('tensorboardX')

        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, name):
                return lambda *a, **k: None

        null.SummaryWriter = SummaryWriter
        sys.modules['tensorboardX'] = null
    return avg

"""Launcher that runs the reference's UNCHANGED scripts (train_rnr.py, test_rnr.py, train_dnr.py, test_dnr.py, precompute.py)
on top of the B200 drop-in modules (SURVEY.md 8b "How the unchanged scripts bind to the replacement"):

    python -m relightable_nr_b200.run /path/to/relightable-nr/train_rnr.py --data_root ... --gpu_id 0
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m relightable_nr_b200.run /path/to/train_rnr.py ... --gpu_id 0   # data parallel

The scripts do plain top-level imports (``import network``, ``import neural_renderer as nr``, train_rnr.py:14-24) and Python
puts the script's own directory first on ``sys.path``, so the drop-ins are registered in ``sys.modules`` BEFORE the script
starts: ``network, render, camera, sph_harm, misc, neural_renderer(.cuda.*), pytorch_prototyping(.pytorch_prototyping),
gcn_lib(.dense/.sparse)`` resolve to relightable_nr_b200/dropin; ``dataio, data_util, util, metric`` (host I/O, out of scope)
keep resolving to the reference's own files next to the script.

Compatibility shims for this software stack (part of the boundary, not of the reference): ``np.int`` / ``np.float`` aliases
(removed in numpy 1.24); ``tensorboardX.SummaryWriter`` -> ``torch.utils.tensorboard`` (or a no-op writer);
``torch_geometric.data.Data`` (attribute bag with ``.to()``: the only thing train_rnr.py:258 uses); ``torch_cluster`` (imported
by gcn_lib, never called on the dense path); ``pytorch_msssim`` (metric.py:3: a Gaussian-window SSIM in relightable_nr_b200/compat), ``skimage.transform``
(data_util.py:3), ``pyshtools`` (replaced by the drop-in sph_harm), ``trimesh``: registered only when the real package is
absent; ``torchvision.utils.make_grid(range=...)`` (renamed ``value_range``); ``OPENCV_IO_ENABLE_OPENEXR=1``.
"""
import importlib
import os
import runpy
import sys
import types


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__rnr_stub__ = True
    sys.modules[name] = m
    return m


class _NullWriter:
    """tensorboardX.SummaryWriter stand-in when no TensorBoard writer is importable: accepts and drops every call."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


class Data:
    """torch_geometric.data.Data as train_rnr.py:258 uses it: named tensors + ``.to(device)``."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device, *a, **k):
        import torch
        return Data(**{k_: (v.to(device, *a, **k) if isinstance(v, torch.Tensor) else v) for k_, v in self.__dict__.items()})


def install_shims():
    os.environ.setdefault('OPENCV_IO_ENABLE_OPENEXR', '1')
    import numpy as np
    for nm, ty in (('int', int), ('float', float), ('bool', bool)):
        if not hasattr(np, nm):
            setattr(np, nm, ty)
    # tensorboardX
    try:
        importlib.import_module('tensorboardX')
    except Exception:
        try:
            from torch.utils.tensorboard import SummaryWriter
        except Exception:
            SummaryWriter = _NullWriter
        _stub('tensorboardX', SummaryWriter=SummaryWriter)
    # torch_geometric / torch_cluster
    tg = _stub('torch_geometric')
    if getattr(tg, '__rnr_stub__', False):
        tg.data = _stub('torch_geometric.data', Data=Data)
        tg.nn = _stub('torch_geometric.nn')
        tg.utils = _stub('torch_geometric.utils')
    _stub('torch_cluster', knn_graph=None)
    _stub('pyshtools')
    # trimesh.load(obj, process=False).vertices / .faces (stitch_lp.py:96-97)
    try:
        importlib.import_module('trimesh')
    except Exception:
        from .compat import trimesh as _trimesh_mod
        sys.modules['trimesh'] = _trimesh_mod
    # pytorch_msssim.ssim: called by the validation pass of the training scripts (metric.py:78-84), which the unchanged
    # train_rnr.py / train_dnr.py execute at iteration 0 -- a real Gaussian-window SSIM, not a stub
    try:
        importlib.import_module('pytorch_msssim')
    except Exception:
        from .compat import pytorch_msssim as _ssim_mod
        sys.modules['pytorch_msssim'] = _ssim_mod
    sk = _stub('skimage')
    if getattr(sk, '__rnr_stub__', False):
        sk.transform = _stub('skimage.transform')
        sk.io = _stub('skimage.io')
    # torchvision.utils.make_grid(range=...) -> value_range
    try:
        import inspect
        import torchvision.utils as tvu
        if 'range' not in inspect.signature(tvu.make_grid).parameters and not getattr(tvu.make_grid, '__rnr_wrapped__', False):
            _orig = tvu.make_grid

            def make_grid(tensor, *a, range=None, **k):
                if range is not None and 'value_range' not in k:
                    k['value_range'] = range
                return _orig(tensor, *a, **k)

            make_grid.__rnr_wrapped__ = True
            tvu.make_grid = make_grid
    except Exception:
        pass


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ('-h', '--help'):
        print(__doc__)
        return 0
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit('relightable_nr_b200.run: no such script: %s' % script)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        _enter_data_parallel()
    install_shims()
    from . import dropin
    installed = dropin.install()
    print('relightable_nr_b200.run: drop-in modules registered: %s' % ', '.join(installed), file=sys.stderr)
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and rank != 0:
        argv = _per_rank_log_dir(argv, rank)
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))          # dataio / data_util / util / metric: the reference's own files
    if world > 1 and rank != 0:
        _mute_checkpoints()
    runpy.run_path(script, run_name='__main__')
    return 0


def _enter_data_parallel():
    """Launched by torchrun (one process per GPU): pin this process to its GPU *before* the script's own
    ``os.environ["CUDA_VISIBLE_DEVICES"] = opt.gpu_id`` (train_rnr.py:148-150) can matter -- CUDA is initialised here, so that
    later write is inert and the script's ``cuda:0`` (pass ``--gpu_id 0``) is this rank's device -- then join the process group and
    install the sampler / gradient-averaging hooks (relightable_nr_b200/parallel.py::install_script_hooks)."""
    local = os.environ.get('LOCAL_RANK', os.environ.get('RANK', '0'))
    os.environ['CUDA_VISIBLE_DEVICES'] = local
    import torch
    import torch.distributed as dist
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.init()
        torch.cuda.set_device(0)
    if not dist.is_initialized():
        dist.init_process_group('nccl' if cuda else 'gloo')
    from . import parallel
    # identical construction-time randomness on every rank (buffers such as spectral-norm u / v and BatchNorm statistics start
    # equal; the parameters are additionally broadcast from rank 0 when the optimizer is built)
    seed = int(os.environ.get('RNR_DP_SEED', '0'))
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    parallel.install_script_hooks(seed=seed)
    print('relightable_nr_b200.run: data parallel, rank %d of %d (%s)' % (dist.get_rank(), dist.get_world_size(),
                                                                          'nccl' if cuda else 'gloo'), file=sys.stderr)


def _per_rank_log_dir(argv, rank):
    """Ranks > 0 still run the script's validation / image dumps (train_rnr.py:626-887): give each its own log directory so
    they do not race on rank 0's files -- the scripts append ``_<exp_name>`` to the directory name (train_rnr.py:446-450)."""
    argv = list(argv)
    tag = 'rank%d' % rank
    if '--exp_name' in argv[:-1]:
        i = argv.index('--exp_name')
        argv[i + 1] = (argv[i + 1] + '_' + tag) if argv[i + 1] else tag
    else:
        argv += ['--exp_name', tag]
    return argv


def _mute_checkpoints():
    """Only rank 0 writes checkpoints (util.custom_save, util.py:33-47); the replicas are identical anyway."""
    try:
        util = importlib.import_module('util')
        if hasattr(util, 'custom_save'):
            util.custom_save = lambda *a, **k: None
    except Exception:
        pass


if __name__ == '__main__':
    sys.exit(main())

"""Import the *real* reference modules from /root/reference under import shims (build container only).

/root/reference does not exist on the GPU box; everything that uses this module is either the golden
generator (tests/golden/make_golden.py) or a test that skips when the directory is absent.
Shims: np.int alias; stub modules for torch_cluster, torch_geometric, pyshtools,
skimage, neural_renderer (the CUDA extension cannot be imported without a GPU build).
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'baseline', '_ref', 'relightable-nr')
# the build container has the reference itself; the GPU box only the copy staged by tools/stage_reference.py (git-ignored)
REF = '/root/reference' if os.path.isdir('/root/reference') else _STAGED


def available():
    return os.path.isdir(REF)


def import_reference():
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int
    if not hasattr(np, 'float'):
        np.float = float

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Data:
        def __init__(self, **kw):
            self.__dict__.update(kw)

        def to(self, *a, **k):
            return self

    stub('torch_cluster', knn_graph=None)
    tg = stub('torch_geometric')
    tg.data = stub('torch_geometric.data', Data=_Data)
    tg.nn = stub('torch_geometric.nn', MessagePassing=object, EdgeConv=object, GCNConv=object, SAGEConv=object,
                 GATConv=object, GINConv=object)
    tg.utils = stub('torch_geometric.utils', remove_self_loops=None, add_self_loops=None)
    stub('pyshtools')
    sk = stub('skimage')
    sk.transform = stub('skimage.transform')
    sk.io = stub('skimage.io')
    nr = stub('neural_renderer')
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    mods = {}
    for name in ('misc', 'camera', 'render', 'sph_harm', 'data_util', 'network'):
        mods[name] = importlib.import_module(name)
    mods['pytorch_prototyping'] = importlib.import_module('pytorch_prototyping.pytorch_prototyping')
    # pyshtools (sph_harm.py:66-68) is not installable here: the reference's evaluate_sh_basis is the ONE function replaced,
    # by the scipy-pinned oracle restatement (tests/test_oracle_golden.py); everything else is the reference's own code
    from oracle import pixel_ops as _P

    def evaluate_sh_basis(lmax=0, azi=None, pol=None, directions=None):
        import numpy as _np
        if directions is None:
            a, p = _np.deg2rad(_np.asarray(azi, dtype=_np.float64)), _np.deg2rad(_np.asarray(pol, dtype=_np.float64))
            directions = _np.stack((_np.sin(p) * _np.cos(a), _np.sin(p) * _np.sin(a), _np.cos(p)), -1)
        return _P.evaluate_sh_basis(lmax, _np.asarray(directions))

    mods['sph_harm'].evaluate_sh_basis = evaluate_sh_basis
    return types.SimpleNamespace(**mods)

"""Import the *real* reference modules from /root/reference under import shims (build container only).

/root/reference does not exist on the GPU box; everything that uses this module is either the golden
generator (tests/golden/make_golden.py) or a test that skips when the directory is absent.
Shims: np.int alias; stub modules for torch_cluster, torch_geometric, pyshtools,
skimage, neural_renderer (the CUDA extension cannot be imported without a GPU build).
"""
import os
import sys
import types

REF = '/root/reference'


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
    return types.SimpleNamespace(**mods)

"""``gcn_lib.sparse`` (torch_geometric edge-list variants) is never imported by the reference's scripts;
the package exists so that ``import gcn_lib.sparse`` resolves, and says so when something is requested from it."""


def __getattr__(name):
    raise AttributeError('gcn_lib.sparse.%s is not part of the per-view hot path (use gcn_lib.dense)' % name)

"""``gcn_lib.sparse`` of the reference (gcn_lib/sparse/{torch_nn,torch_edge,torch_vertex}.py): the edge-list ("sparse", torch_geometric
style) flavour of the graph convolutions -- node features x [V, C], edge_index [2, E] = (neighbour, centre).  network.DenseDeepGCN
only uses the dense flavour; these are the same class names over the same kernels (north_star: "gcn_lib/dense+sparse")."""
from .torch_nn import *       # noqa: F401,F403
from .torch_edge import *     # noqa: F401,F403
from .torch_vertex import *   # noqa: F401,F403

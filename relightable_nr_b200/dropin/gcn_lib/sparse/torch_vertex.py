"""gcn_lib/sparse/torch_vertex.py of the reference: graph convolutions on node-major features x [V, C] with an edge list.

``EdgConv`` (torch_geometric.nn.EdgeConv(MLP([2C, C']), aggr='max') in the reference, torch_vertex.py:23-31) is the same operator
as the dense ``EdgeConv4D``: max over a node's incoming edges of MLP(cat[x_i, x_j - x_i]).  When the edge list is the regular
k-nearest-neighbour list the DilatedKnnGraph produces (every node has k incoming edges, stored node after node) and the MLP is
Linear -> (ReLU | LeakyReLU) -> BatchNorm1d, it runs on the fused P|Q + gather / max kernels of csrc/gcn.cu through the dense
operator; any other edge list goes through plain scatter reductions."""
import torch
from torch import nn

from ..dense.torch_vertex import _EdgeConvFn, _fusable
from .torch_edge import DilatedKnnGraph
from .torch_nn import MLP

__all__ = ['MRConv', 'EdgConv', 'GraphConv', 'DynConv', 'ResDynBlock', 'DenseDynBlock']


def _scatter(aggr, src, index, num_nodes):
    out = torch.zeros((num_nodes, src.shape[1]), dtype=src.dtype, device=src.device)
    idx = index.view(-1, 1).expand_as(src)
    if aggr == 'max':
        return out.scatter_reduce(0, idx, src, reduce='amax', include_self=False)
    if aggr == 'mean':
        return out.scatter_reduce(0, idx, src, reduce='mean', include_self=False)
    if aggr in ('add', 'sum'):
        return out.scatter_add(0, idx, src)
    raise NotImplementedError('aggr %s' % aggr)


def _regular_knn(edge_index, V):
    """k when edge_index [2, V*k] lists exactly k incoming edges per node, node after node (centre = 0,0,..,1,1,..); else 0."""
    E = edge_index.shape[1]
    if V == 0 or E % V:
        return 0
    k = E // V
    centre = torch.arange(V, device=edge_index.device).repeat_interleave(k)
    return k if torch.equal(edge_index[1], centre) else 0


class MRConv(nn.Module):
    """Max-relative graph convolution (torch_vertex.py:8-20)."""

    def __init__(self, in_channels, out_channels, act_type='relu', norm_type=None, bias=True, aggr='max'):
        super().__init__()
        self.nn = MLP([in_channels * 2, out_channels], act_type, norm_type, bias)
        self.aggr = aggr

    def forward(self, x, edge_index):
        rel = _scatter(self.aggr, x.index_select(0, edge_index[0]) - x.index_select(0, edge_index[1]), edge_index[1], x.shape[0])
        return self.nn(torch.cat([x, rel], dim=1))


class EdgConv(nn.Module):
    """Edge convolution (torch_vertex.py:23-31); parameters live under ``nn`` like torch_geometric's EdgeConv."""

    def __init__(self, in_channels, out_channels, act_type='relu', norm_type=None, bias=True, aggr='max'):
        super().__init__()
        self.nn = MLP([in_channels * 2, out_channels], act_type, norm_type, bias)
        self.aggr = aggr

    def forward(self, x, edge_index):
        V = x.shape[0]
        k = _regular_knn(edge_index, V) if (self.aggr == 'max' and x.is_cuda) else 0
        if k:
            x4 = x.t()[None, :, :, None]
            ei = torch.stack((edge_index[0].view(1, V, k), edge_index[1].view(1, V, k)), 0)
            if _fusable(self.nn, x4, ei) is not None:
                out = _EdgeConvFn.apply(self.nn, ei, None, x4, *self.nn.parameters())
                return out[0, :, :, 0].t()
        x_i, x_j = x.index_select(0, edge_index[1]), x.index_select(0, edge_index[0])
        return _scatter(self.aggr, self.nn(torch.cat([x_i, x_j - x_i], dim=1)), edge_index[1], V)


class GraphConv(nn.Module):
    """Static graph convolution layer (torch_vertex.py:34-47)."""

    def __init__(self, in_channels, out_channels, conv_type='edge', act_type='relu', norm_type=None, bias=True):
        super().__init__()
        if conv_type == 'edge':
            self.gconv = EdgConv(in_channels, out_channels, act_type, norm_type, bias)
        elif conv_type == 'mr':
            self.gconv = MRConv(in_channels, out_channels, act_type, norm_type, bias)
        else:
            raise NotImplementedError('conv_type is not supported')

    def forward(self, x, edge_index):
        return self.gconv(x, edge_index)


class DynConv(GraphConv):
    """Dynamic graph convolution: the kNN graph is rebuilt from the input features (torch_vertex.py:50-63)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv_type='edge', act_type='relu', norm_type=None, bias=True,
                 stochastic=False, epsilon=1.0, knn_type='matrix'):
        super().__init__(in_channels, out_channels, conv_type, act_type, norm_type, bias)
        self.k, self.d = kernel_size, dilation
        self.dilated_knn_graph = DilatedKnnGraph(kernel_size, dilation, stochastic, epsilon, knn_type)

    def forward(self, x, batch=None):
        return super().forward(x, self.dilated_knn_graph(x, batch))


class ResDynBlock(nn.Module):
    """(x, batch) -> (DynConv(x) + x, batch) (torch_vertex.py:66-80)."""

    def __init__(self, channels, kernel_size=9, dilation=1, conv_type='edge', act_type='relu', norm_type=None, bias=True, stochastic=False,
                 epsilon=1.0, knn_type='matrix'):
        super().__init__()
        self.body = DynConv(channels, channels, kernel_size, dilation, conv_type, act_type, norm_type, bias, stochastic, epsilon, knn_type)

    def forward(self, x, batch):
        return self.body(x, batch) + x, batch


class DenseDynBlock(nn.Module):
    """(x, batch) -> (cat(x, DynConv(x)), batch).  (The reference's forward calls ``self.body(batch)`` and its body is built for
    2*channels inputs, torch_vertex.py:83-95 -- it cannot run as written; this is what it evidently intends.)"""

    def __init__(self, channels, kernel_size=9, dilation=1, conv_type='edge', act_type='relu', norm_type=None, bias=True, stochastic=False,
                 epsilon=1.0, knn_type='matrix'):
        super().__init__()
        self.body = DynConv(channels, channels, kernel_size, dilation, conv_type, act_type, norm_type, bias, stochastic, epsilon, knn_type)

    def forward(self, x, batch):
        return torch.cat((x, self.body(x, batch)), 1), batch

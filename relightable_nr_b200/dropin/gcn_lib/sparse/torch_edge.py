"""gcn_lib/sparse/torch_edge.py of the reference: dilated kNN graphs as edge lists [2, V*k] = (neighbour, centre)."""
import torch
from torch import nn

__all__ = ['Dilated', 'DilatedKnnGraph', 'pairwise_distance', 'knn_matrix', 'knn_graph_matrix']


class Dilated(nn.Module):
    """Every ``dilation``-th of the k*dilation neighbours of each node -- or, with probability ``epsilon`` while training and
    ``stochastic``, a random k of them (torch_edge.py:6-29)."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k

    def forward(self, edge_index, batch=None):
        if self.stochastic and torch.rand(1) < self.epsilon and self.training:
            num = self.k * self.dilation
            pick = torch.randperm(num)[:self.k].to(edge_index.device)
            return edge_index.view(2, -1, num)[:, :, pick].reshape(2, -1)
        return edge_index[:, ::self.dilation]


def pairwise_distance(x):
    """x [B, V, C] -> squared distances [B, V, V] (torch_edge.py:53-63)."""
    inner = -2 * torch.matmul(x, x.transpose(2, 1))
    sq = torch.sum(x * x, dim=-1, keepdim=True)
    return sq + inner + sq.transpose(2, 1)


def knn_matrix(x, k=16, batch=None):
    """x [B*V, C], batch [B*V] (graph id of each node, equal-sized graphs) -> (neighbour [1, B*V*k], centre [1, B*V*k]) (torch_edge.py:66-95)."""
    B = int(batch[-1]) + 1 if batch is not None else 1
    x = x.view(B, -1, x.shape[-1])
    V = x.shape[1]
    nn_idx = torch.topk(-pairwise_distance(x), k=k)[1]
    nn_idx = nn_idx + torch.arange(0, V * B, V, device=x.device).view(B, 1, 1)
    center = torch.arange(0, V * B, device=x.device).repeat(k, 1).transpose(1, 0).contiguous().view(1, -1)
    return nn_idx.view(1, -1), center


def knn_graph_matrix(x, k=16, batch=None):
    nn_idx, center = knn_matrix(x, k, batch)
    return torch.cat((nn_idx, center), dim=0)


class DilatedKnnGraph(nn.Module):
    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0, knn_type='matrix'):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k
        self._dilated = Dilated(k, dilation, stochastic, epsilon)
        if knn_type == 'matrix':
            self.knn = knn_graph_matrix
        else:
            try:
                from torch_cluster import knn_graph
            except Exception:
                knn_graph = None
            if knn_graph is None:
                raise NotImplementedError("knn_type != 'matrix' needs torch_cluster.knn_graph, which is not installed")
            self.knn = knn_graph

    def forward(self, x, batch):
        return self._dilated(self.knn(x, self.k * self.dilation, batch), batch)

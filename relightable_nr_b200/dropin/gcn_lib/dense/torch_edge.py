"""gcn_lib/dense/torch_edge.py of the reference: dense dilated kNN graphs.  The V x V distance matrix is one GEMM and the
neighbour selection one top-k (library calls: this stage is off the per-view critical path)."""
import torch
from torch import nn

__all__ = ['DenseDilated', 'pairwise_distance', 'dense_knn_matrix', 'DenseDilatedKnnGraph']


class DenseDilated(nn.Module):
    """Keep every ``dilation``-th of the k*dilation nearest neighbours -- or, with probability ``epsilon`` while training and
    ``stochastic``, a random k of them (drawn from the CPU generator like the reference, torch_edge.py:19-29).
    edge_index: [2, B, V, k*dilation]."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k

    def forward(self, edge_index):
        if self.stochastic and torch.rand(1) < self.epsilon and self.training:
            pick = torch.randperm(self.k * self.dilation)[:self.k]
            return edge_index[:, :, :, pick.to(edge_index.device)]
        return edge_index[:, :, :, ::self.dilation]


def pairwise_distance(x):
    """x [B,V,C] -> squared distances [B,V,V] (torch_edge.py:32-43)."""
    inner = -2 * torch.matmul(x, x.transpose(2, 1))
    sq = torch.sum(x * x, dim=-1, keepdim=True)
    return sq + inner + sq.transpose(2, 1)


def dense_knn_matrix(x, k=16):
    """x [B,V,C,1] -> edge_index [2,B,V,k] = (neighbour, centre) (torch_edge.py:46-65)."""
    x = x.squeeze(-1)
    B, V, _ = x.shape
    nn_idx = torch.topk(-pairwise_distance(x), k=k)[1]
    center = torch.arange(V, device=x.device).view(1, V, 1).expand(B, V, k)
    return torch.stack((nn_idx, center), dim=0)


class DenseDilatedKnnGraph(nn.Module):
    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)
        self.knn = dense_knn_matrix

    def forward(self, x):
        return self._dilated(self.knn(x, self.k * self.dilation))

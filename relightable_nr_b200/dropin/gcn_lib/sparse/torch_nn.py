"""gcn_lib/sparse/torch_nn.py of the reference: layer factories on [num_nodes, C] features (Linear / BatchNorm1d)."""
from torch import nn
from torch.nn import Sequential as Seq, Linear as Lin

__all__ = ['act_layer', 'norm_layer', 'MultiSeq', 'MLP']


def act_layer(act_type, inplace=False, neg_slope=0.2, n_prelu=1):
    kind = act_type.lower()
    if kind == 'relu':
        return nn.ReLU(inplace)
    if kind == 'leakyrelu':
        return nn.LeakyReLU(neg_slope, inplace)
    if kind == 'prelu':
        return nn.PReLU(num_parameters=n_prelu, init=neg_slope)
    raise NotImplementedError('activation layer [%s] is not found' % kind)


def norm_layer(norm_type, nc):
    kind = norm_type.lower()
    if kind == 'batch':
        return nn.BatchNorm1d(nc, affine=True)
    if kind == 'instance':
        return nn.InstanceNorm1d(nc, affine=False)
    raise NotImplementedError('normalization layer [%s] is not found' % kind)


class MultiSeq(Seq):
    """Sequential whose stages may take / return tuples (torch_nn.py:42-52)."""

    def forward(self, *inputs):
        for module in self._modules.values():
            inputs = module(*inputs) if type(inputs) == tuple else module(inputs)
        return inputs


class MLP(Seq):
    """Linear -> act -> norm per stage (torch_nn.py:55-64: activation BEFORE the norm, norm sized by the LAST width)."""

    def __init__(self, channels, act_type='relu', norm_type=None, bias=True):
        m = []
        for cin, cout in zip(channels[:-1], channels[1:]):
            m.append(Lin(cin, cout, bias))
            if act_type:
                m.append(act_layer(act_type))
            if norm_type:
                m.append(norm_layer(norm_type, channels[-1]))
        super().__init__(*m)

"""gcn_lib/dense/torch_nn.py of the reference: layer factories with the reference's names and module layouts (state-dict keys
``<seq>.0.weight`` for the 1x1 convolution, ``<seq>.2.*`` for the normalisation: activation sits BEFORE the norm, torch_nn.py:55-64)."""
import torch
from torch import nn
from torch.nn import Sequential as Seq, Linear as Lin, Conv2d

__all__ = ['act_layer', 'norm_layer', 'MLP', 'BasicConv', 'batched_index_select']


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
        return nn.BatchNorm2d(nc, affine=True)
    if kind == 'instance':
        return nn.InstanceNorm2d(nc, affine=False)
    raise NotImplementedError('normalization layer [%s] is not found' % kind)


def _stack(make_linear, channels, act_type, norm_type):
    layers = []
    for cin, cout in zip(channels[:-1], channels[1:]):
        layers.append(make_linear(cin, cout))
        if act_type:
            layers.append(act_layer(act_type))
        if norm_type:
            layers.append(norm_layer(norm_type, channels[-1]))
    return layers


class MLP(Seq):
    def __init__(self, channels, act_type='relu', norm_type=None, bias=True):
        super().__init__(*_stack(lambda a, b: Lin(a, b, bias), channels, act_type, norm_type))


class BasicConv(Seq):
    def __init__(self, channels, act_type='relu', norm_type=None, bias=True):
        super().__init__(*_stack(lambda a, b: Conv2d(a, b, 1, bias=bias), channels, act_type, norm_type))


def batched_index_select(inputs, index):
    """inputs [B,C,V,1], index [B,V,k] -> [B,C,V,k] (torch_nn.py:70-85)."""
    B, C, V, _ = inputs.shape
    k = index.shape[2]
    flat = inputs[..., 0].permute(0, 2, 1).reshape(B * V, C)
    idx = (index + torch.arange(B, device=index.device, dtype=index.dtype).view(B, 1, 1) * V).reshape(-1)
    return flat.index_select(0, idx).view(B, V * k, C).permute(0, 2, 1).reshape(B, C, V, k)

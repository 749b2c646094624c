"""gcn_lib/dense/torch_vertex.py of the reference: graph convolutions on dense [B,C,V,1] features.

``EdgeConv4D`` runs on librnr_b200's fused kernels (csrc/gcn.cu) for CUDA inputs with batch 1 -- the configuration
network.DenseDeepGCN uses: one GEMM for P|Q, one gather/activation/max-min/statistics pass, one BatchNorm finish.  Autograd is
supported by recomputing the operator with torch ops in backward (the reference never back-propagates into the GCN: its output is
dead).  Other inputs (CPU tensors at construction time, batch > 1, PReLU / InstanceNorm) use the torch-op form."""
import ctypes as C

import torch
from torch import nn

from .... import _lib
from .torch_edge import DenseDilatedKnnGraph
from .torch_nn import BasicConv, batched_index_select

__all__ = ['MRConv4D', 'EdgeConv4D', 'GraphConv4D', 'DynConv4D', 'ResDynBlock4D', 'DenseDynBlock4D']

vp, i32, f32, f64 = C.c_void_p, C.c_int, C.c_float, C.c_double
_lib.register_sigs({
    "rnr_edgeconv_reduce": [vp, vp, i32, i32, i32, f32, vp, vp, vp, vp],
    "rnr_edgeconv_finish": [vp, vp, vp, f64, vp, vp, f32, vp, vp, f32, i32, vp, vp, i32, i32, vp],
})


def _edgeconv_torch(seq, x, edge_index):
    x_i = batched_index_select(x, edge_index[1])
    x_j = batched_index_select(x, edge_index[0])
    feat = torch.cat([x_i, x_j - x_i], dim=1)                         # [B, 2C, V, k]
    if isinstance(seq[0], nn.Linear):
        # node-major MLP of gcn_lib.sparse (Linear / BatchNorm1d over the E = V*k edge rows): same numbers, other layout
        B, C2, V, K = feat.shape
        out = seq(feat.permute(0, 2, 3, 1).reshape(B * V * K, C2)).view(B, V, K, -1).permute(0, 3, 1, 2)
        return torch.max(out, -1, keepdim=True)[0]
    return torch.max(seq(feat), -1, keepdim=True)[0]


def _fusable(seq, x, edge_index):
    if not (x.is_cuda and x.dtype == torch.float32 and x.shape[0] == 1 and x.shape[-1] == 1):
        return None
    mods = list(seq)
    if not mods:
        return None
    if isinstance(mods[0], nn.Conv2d):
        if mods[0].kernel_size != (1, 1):
            return None
    elif not isinstance(mods[0], nn.Linear):           # (nn.Linear: the node-major MLP of gcn_lib.sparse -- the same 1x1 map)
        return None
    conv, act, bn = mods[0], None, None
    for m in mods[1:]:
        if isinstance(m, (nn.ReLU, nn.LeakyReLU)) and act is None and bn is None:
            act = m
        elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)) and bn is None:
            bn = m
        else:
            return None
    if bn is not None and not bn.affine:
        return None
    slope = 1.0 if act is None else (0.0 if isinstance(act, nn.ReLU) else float(act.negative_slope))
    return conv, bn, slope


class _EdgeConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seq, edge_index, residual, x, *params):
        conv, bn, slope = _fusable(seq, x, edge_index)
        L = _lib.lib()
        s = torch.cuda.current_stream().cuda_stream
        V, K = edge_index.shape[2], edge_index.shape[3]
        Cin, Cout = x.shape[1], conv.weight.shape[0]
        X = x[0, :, :, 0].t().contiguous()                                  # [V, Cin]
        W = conv.weight.detach().reshape(Cout, -1)                          # [Cout, 2 Cin] acting on cat[x_i, x_j - x_i]
        Wcat = torch.cat((W[:, :Cin] - W[:, Cin:], W[:, Cin:]), 0).t().contiguous()      # [Cin, 2 Cout] -> P | Q
        bias = torch.zeros(2 * Cout, dtype=torch.float32, device=x.device)
        if conv.bias is not None:
            bias[:Cout] = conv.bias.detach()
        pq = torch.addmm(bias, X, Wcat)                                     # [V, 2 Cout]   (plain library GEMM)
        nbr = edge_index[0, 0].to(torch.int32).contiguous()
        amax = torch.empty((V, Cout), dtype=torch.float32, device=x.device)
        amin = torch.empty_like(amax)
        training = bn is not None and (bn.training or bn.running_mean is None)
        sums = torch.zeros(2 * Cout, dtype=torch.float64, device=x.device) if training else None
        _lib.check(L.rnr_edgeconv_reduce(pq.data_ptr(), nbr.data_ptr(), V, K, Cout, slope, amax.data_ptr(), amin.data_ptr(),
                                         sums.data_ptr() if sums is not None else None, s), 'rnr_edgeconv_reduce')
        out = torch.empty((V, Cout), dtype=torch.float32, device=x.device)
        res = residual[0, :, :, 0].t().contiguous() if residual is not None else None
        track = bn is not None and bn.training and bn.track_running_stats
        mom = 0.1 if (bn is None or bn.momentum is None) else float(bn.momentum)
        _lib.check(L.rnr_edgeconv_finish(
            amax.data_ptr(), amin.data_ptr(), sums.data_ptr() if sums is not None else None, float(V * K),
            bn.weight.data_ptr() if bn is not None else None, bn.bias.data_ptr() if bn is not None else None,
            float(bn.eps) if bn is not None else 0.0,
            bn.running_mean.data_ptr() if (bn is not None and bn.running_mean is not None and (track or not training)) else None,
            bn.running_var.data_ptr() if (bn is not None and bn.running_var is not None and (track or not training)) else None,
            mom, 1 if training else 0, res.data_ptr() if res is not None else None, out.data_ptr(), V, Cout, s), 'rnr_edgeconv_finish')
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        ctx.seq, ctx.edge_index, ctx.has_res = seq, edge_index, residual is not None
        ctx.save_for_backward(x)
        return out.t()[None, :, :, None]

    @staticmethod
    def backward(ctx, g):
        # recompute with torch ops (batch statistics re-derived; running statistics left alone) and differentiate that
        (x,) = ctx.saved_tensors
        seq = ctx.seq
        bns = [m for m in seq if isinstance(m, nn.BatchNorm2d)]
        saved = [(m.momentum, m.track_running_stats) for m in bns]
        with torch.enable_grad():
            xx = x.detach().requires_grad_(True)
            for m in bns:
                m.momentum = 0.0
            y = _edgeconv_torch(seq, xx, ctx.edge_index)
            for m, (mo, _) in zip(bns, saved):
                m.momentum = mo
            params = [p for p in seq.parameters()]
            grads = torch.autograd.grad(y, [xx] + params, g, allow_unused=True)
        gx = grads[0] + (g if ctx.has_res else 0)
        return (None, None, g if ctx.has_res else None, gx, *grads[1:])


class MRConv4D(nn.Module):
    """Max-relative graph convolution (torch_vertex.py:8-20)."""

    def __init__(self, in_channels, out_channels, act_type='relu', norm_type=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act_type, norm_type, bias)

    def forward(self, x, edge_index):
        x_i = batched_index_select(x, edge_index[1])
        x_j = batched_index_select(x, edge_index[0])
        rel = torch.max(x_j - x_i, -1, keepdim=True)[0]
        return self.nn(torch.cat([x, rel], dim=1))


class EdgeConv4D(nn.Module):
    """Edge convolution: max over the k neighbours of nn(cat[x_i, x_j - x_i]) (torch_vertex.py:23-35)."""

    def __init__(self, in_channels, out_channels, act_type='relu', norm_type=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act_type, norm_type, bias)

    def forward(self, x, edge_index, residual=None):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise TypeError('EdgeConv4D input must be a CUDA tensor (librnr_b200 has no CPU path)')
        if _fusable(self.nn, x, edge_index) is not None:
            return _EdgeConvFn.apply(self.nn, edge_index, residual, x, *self.nn.parameters())
        out = _edgeconv_torch(self.nn, x, edge_index)
        return out if residual is None else out + residual


class GraphConv4D(nn.Module):
    """Static graph convolution layer (torch_vertex.py:38-53)."""

    def __init__(self, in_channels, out_channels, conv_type='edge', act_type='relu', norm_type=None, bias=True):
        super().__init__()
        if conv_type == 'edge':
            self.gconv = EdgeConv4D(in_channels, out_channels, act_type, norm_type, bias)
        elif conv_type == 'mr':
            self.gconv = MRConv4D(in_channels, out_channels, act_type, norm_type, bias)
        else:
            raise NotImplementedError('conv_type is not supported')

    def forward(self, x, edge_index):
        return self.gconv(x, edge_index)


class DynConv4D(GraphConv4D):
    """Dynamic graph convolution: the kNN graph is rebuilt from the input features (torch_vertex.py:56-70)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv_type='edge', act_type='relu', norm_type=None,
                 bias=True, stochastic=False, epsilon=0.0):
        super().__init__(in_channels, out_channels, conv_type, act_type, norm_type, bias)
        self.k, self.d = kernel_size, dilation
        self.dilated_knn_graph = DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)

    def forward(self, x, residual=None):
        edge_index = self.dilated_knn_graph(x.transpose(2, 1))
        if residual is not None and isinstance(self.gconv, EdgeConv4D):
            return self.gconv(x, edge_index, residual=residual)            # residual add fused into the finish kernel
        out = self.gconv(x, edge_index)
        return out if residual is None else out + residual


class ResDynBlock4D(nn.Module):
    """x + DynConv(x) (torch_vertex.py:73-86)."""

    def __init__(self, channels, kernel_size=9, dilation=1, conv_type='edge', act_type='relu', norm_type=None, bias=True,
                 stochastic=False, epsilon=0.0):
        super().__init__()
        self.body = DynConv4D(channels, channels, kernel_size, dilation, conv_type, act_type, norm_type, bias, stochastic, epsilon)

    def forward(self, x):
        return self.body(x, residual=x)


class DenseDynBlock4D(nn.Module):
    """cat(x, DynConv(x)) (torch_vertex.py:89-102)."""

    def __init__(self, in_channels, out_channels=64, kernel_size=9, dilation=1, conv_type='edge', act_type='relu', norm_type=None,
                 bias=True, stochastic=False, epsilon=0.0):
        super().__init__()
        self.body = DynConv4D(in_channels, out_channels, kernel_size, dilation, conv_type, act_type, norm_type, bias, stochastic, epsilon)

    def forward(self, x):
        return torch.cat((x, self.body(x)), 1)

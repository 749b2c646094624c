"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of network.DenseDeepGCN (/root/reference/network.py:256-315) and the
gcn_lib/dense operators it uses: dense kNN graph (gcn_lib/dense/torch_edge.py:32-83), EdgeConv with activation BEFORE
normalisation (torch_vertex.py:23-35, torch_nn.py:55-64), residual dynamic blocks (torch_vertex.py:68-83), channel-max fusion and
the two spectral-normalised Linear layers (torch.nn.utils.spectral_norm, one power iteration per training forward).
Pinned by tests/golden/gcn_small.npz (generated from the real reference module, tests/golden/make_golden.py).
Used by tests/ only; never by the product path."""
import torch
import torch.nn.functional as F


def knn_neighbours(x, k):
    """x [V,C] -> [V,k] indices of the k nearest rows (self included), ordered like torch.topk of -distance (torch_edge.py:32-57)."""
    inner = -2 * (x @ x.t())
    sq = (x * x).sum(-1, keepdim=True)
    neg = -(sq + inner + sq.t())
    return torch.topk(neg, k=k)[1]


def dilate(nn_idx, k, d, stochastic=False, epsilon=0.0, training=False):
    """torch_edge.py:19-29 (the random branch draws from the CPU generator: torch.rand(1), torch.randperm)."""
    if stochastic and torch.rand(1) < epsilon and training:
        return nn_idx[:, torch.randperm(k * d)[:k]]
    return nn_idx[:, ::d]


def edge_conv(x, nbr, w, b, bn=None, slope=0.0, training=True, eps=1e-5):
    """x [V,C]; nbr [V,k]; w [Cout, 2C]; bn = (gamma, beta, running_mean, running_var) or None -> [V,Cout].
    conv1x1(cat[x_i, x_j - x_i]) -> act -> BatchNorm2d (batch statistics over V*k when training) -> max over k."""
    xi = x[:, None, :].expand(-1, nbr.shape[1], -1)
    xj = x[nbr]
    y = torch.cat((xi, xj - xi), -1) @ w.t()
    if b is not None:
        y = y + b
    y = torch.where(y > 0, y, y * slope)
    if bn is not None:
        g, be, rm, rv = bn
        if training:
            m = y.mean((0, 1))
            v = y.var((0, 1), unbiased=False)
        else:
            m, v = rm, rv
        y = (y - m) / torch.sqrt(v + eps) * g + be
    return y.max(1)[0]


def spectral_linear(x, w_orig, bias, u, v, training=True, eps=1e-12):
    """torch.nn.utils.spectral_norm(nn.Linear) forward: one power iteration on (u, v) when training, weight / sigma."""
    if training:
        v = F.normalize(w_orig.t() @ u, dim=0, eps=eps)
        u = F.normalize(w_orig @ v, dim=0, eps=eps)
    sigma = torch.dot(u, w_orig @ v)
    return x @ (w_orig / sigma).t() + bias


def dense_deep_gcn_forward(sd, pos, x, n_blocks, k, training=True, slope=0.0, norm=True, stochastic=False, epsilon=0.0):
    """sd: state dict of the reference module (taken BEFORE the forward); pos, x [V,3] -> [1, out_channels_gcn]."""
    def bn(prefix):
        if not norm:
            return None
        return (sd[prefix + '.weight'], sd[prefix + '.bias'], sd[prefix + '.running_mean'], sd[prefix + '.running_var'])

    data = torch.cat((pos, x), 1)
    nbr = dilate(knn_neighbours(data[:, 0:3], k), k, 1, stochastic, epsilon, training)
    f = edge_conv(data, nbr, sd['head.gconv.nn.0.weight'][:, :, 0, 0], sd.get('head.gconv.nn.0.bias'), bn('head.gconv.nn.2'),
                  slope, training)
    feats = [f]
    for i in range(n_blocks - 1):
        d = 1 + i
        nbr = dilate(knn_neighbours(feats[-1], k * d), k, d, stochastic, epsilon, training)
        p = 'backbone.%d.body.gconv.nn' % i
        feats.append(edge_conv(feats[-1], nbr, sd[p + '.0.weight'][:, :, 0, 0], sd.get(p + '.0.bias'), bn(p + '.2'), slope,
                               training) + feats[-1])
    cat = torch.cat(feats, 1)
    fus = cat @ sd['fusion_block.0.weight'][:, :, 0, 0].t() + sd['fusion_block.0.bias']
    fus = torch.where(fus > 0, fus, fus * slope).max(1)[0]                       # max over CHANNELS -> [V]
    h = spectral_linear(fus, sd['linear.0.weight_orig'], sd['linear.0.bias'], sd['linear.0.weight_u'], sd['linear.0.weight_v'], training)
    h = spectral_linear(h, sd['linear.1.weight_orig'], sd['linear.1.bias'], sd['linear.1.weight_u'], sd['linear.1.weight_v'], training)
    return h[None]

"""network.DenseDeepGCN / gcn_lib.dense drop-ins (a21) on the GPU against the reference golden vector and the CPU oracle.

Tolerances: fused EdgeConv vs the torch-op form of the same operator: max-abs <= 1e-4 (fp32; the fused form evaluates
(W1-W2) x_i + W2 x_j instead of W1 x_i + W2 (x_j - x_i)); whole network at the golden size: <= 2e-3 of the output range (four
kNN graphs rebuilt from features that differ by ~1e-6 select the same neighbours for generic inputs)."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_torch_reference():
    """The torch-op form used as the comparison runs nn.Conv2d through cuDNN, which defaults to TF32 (1e-3 errors): force fp32."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _opt(**kw):
    o = dict(n_filters=16, kernel_size=4, act_type='relu', norm_type='batch', bias=True, epsilon=0.0, stochastic=False,
             conv_type='edge', n_blocks=4, num_v_gcn=96, out_channels_gcn=8, in_channels=6, block_type='res')
    o.update(kw)
    return types.SimpleNamespace(**o)


def test_gcn_matches_reference_golden():
    from relightable_nr_b200.dropin import network
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'gcn_small.npz'))
    sd = {k[len('gcn_sd__'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('gcn_sd__')}
    net = network.DenseDeepGCN(_opt())
    assert sorted(net.state_dict().keys()) == sorted(z['gcn_keys'].tolist())
    net.load_state_dict(sd, strict=True)
    net.cuda().train()
    v = torch.from_numpy(z['gcn_v']).cuda()
    out = net(types.SimpleNamespace(pos=v, x=v))
    ref = torch.from_numpy(z['gcn_fea'])
    err = (out.detach().cpu() - ref).abs().max().item()
    print('GCN vs reference golden: max abs %.3e (range %.2f)' % (err, ref.abs().max().item()))
    assert err <= 2e-3 * ref.abs().max().item()
    # BatchNorm bookkeeping of the fused path: running statistics moved, counters bumped
    assert int(net.head.gconv.nn[2].num_batches_tracked) == 1
    assert not torch.equal(net.head.gconv.nn[2].running_mean.cpu(), sd['head.gconv.nn.2.running_mean'])


@pytest.mark.parametrize('act,norm,training', [('relu', 'batch', True), ('leakyrelu', 'batch', True), ('relu', None, True),
                                               ('relu', 'batch', False)])
def test_edgeconv_fused_matches_torch_ops(act, norm, training):
    from relightable_nr_b200.dropin.gcn_lib.dense import EdgeConv4D, dense_knn_matrix
    from relightable_nr_b200.dropin.gcn_lib.dense.torch_vertex import _edgeconv_torch
    import copy
    torch.manual_seed(1)
    V, Cin, Cout, K = 500, 24, 40, 9
    m = EdgeConv4D(Cin, Cout, act, norm, True).cuda()
    if norm:
        with torch.no_grad():
            m.nn[2].weight.copy_(torch.randn(Cout))          # negative scales exercise the min branch
            m.nn[2].running_mean.normal_()
            m.nn[2].running_var.uniform_(0.5, 2.0)
    m.train(training)
    ref_m = copy.deepcopy(m)
    x = torch.randn(1, Cin, V, 1, device='cuda')
    ei = dense_knn_matrix(x.transpose(2, 1), K)
    res = torch.randn(1, Cout, V, 1, device='cuda')
    out = m(x, ei, residual=res)
    ref = _edgeconv_torch(ref_m.nn, x, ei) + res
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 1e-4
    if norm and training:
        assert torch.allclose(m.nn[2].running_mean, ref_m.nn[2].running_mean, atol=1e-5)
        assert torch.allclose(m.nn[2].running_var, ref_m.nn[2].running_var, atol=1e-5)


def test_edgeconv_backward_by_recompute():
    from relightable_nr_b200.dropin.gcn_lib.dense import EdgeConv4D, dense_knn_matrix
    from relightable_nr_b200.dropin.gcn_lib.dense.torch_vertex import _edgeconv_torch
    import copy
    torch.manual_seed(2)
    V, Cin, Cout, K = 200, 8, 16, 5
    m = EdgeConv4D(Cin, Cout, 'relu', 'batch', True).cuda().train()
    ref_m = copy.deepcopy(m)
    x = torch.randn(1, Cin, V, 1, device='cuda', requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    ei = dense_knn_matrix(x.detach().transpose(2, 1), K)
    w = torch.randn(1, Cout, V, 1, device='cuda')
    (m(x, ei) * w).sum().backward()
    (_edgeconv_torch(ref_m.nn, xr, ei) * w).sum().backward()
    assert torch.allclose(x.grad, xr.grad, atol=1e-4, rtol=1e-3)
    for p, q in zip(m.parameters(), ref_m.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-3, rtol=1e-3)


def test_default_size_gcn_runs():
    """train_rnr.py defaults (train_rnr.py:84-95): 7500 vertices, 20 blocks, k = 16, stochastic dilation -> [1, 512]."""
    import time
    from relightable_nr_b200.dropin import network
    opt = _opt(n_filters=64, kernel_size=16, n_blocks=20, num_v_gcn=7500, out_channels_gcn=512, epsilon=0.2, stochastic=True)
    net = network.DenseDeepGCN(opt).cuda().train()
    assert sum(p.numel() for p in net.parameters()) > 18e6
    g = torch.Generator().manual_seed(0)
    v = torch.nn.functional.normalize(torch.randn(7500, 3, generator=g), dim=-1).cuda()
    inp = types.SimpleNamespace(pos=v, x=v)
    with torch.no_grad():
        net(inp)
        torch.cuda.synchronize()
        t0 = time.time()
        out = net(inp)
        torch.cuda.synchronize()
    print('DenseDeepGCN forward, V=7500, 20 blocks: %.1f ms' % ((time.time() - t0) * 1e3))
    assert out.shape == (1, 512) and torch.isfinite(out).all()


def test_sparse_edgconv_runs_on_the_fused_kernels_and_matches_the_scatter_form():
    """gcn_lib.sparse.EdgConv (torch_geometric.nn.EdgeConv(MLP, 'max') in the reference, gcn_lib/sparse/torch_vertex.py:23-31): on the
    regular kNN edge list it runs on the fused P|Q + gather / max kernels; the same module on a shuffled (irregular) edge list goes
    through plain scatter reductions -- same output (1e-4), same BatchNorm running statistics."""
    import copy
    from relightable_nr_b200 import _lib
    from relightable_nr_b200.dropin.gcn_lib import sparse
    torch.manual_seed(0)
    conv = sparse.EdgConv(16, 32, 'relu', 'batch', True).cuda().train()
    twin = copy.deepcopy(conv)
    x = torch.randn(300, 16, device='cuda')
    batch = torch.zeros(300, dtype=torch.long, device='cuda')
    ei = sparse.knn_graph_matrix(x, 8, batch)
    n0 = _lib.lib().rnr_launch_count()
    out = conv(x, ei)
    assert _lib.lib().rnr_launch_count() - n0 >= 2, 'the regular edge list must run on librnr_b200 kernels'
    perm = torch.randperm(ei.shape[1], device='cuda')
    ref = twin(x, ei[:, perm].contiguous())
    assert out.shape == ref.shape == (300, 32)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4), (out - ref).abs().max().item()
    assert torch.allclose(conv.nn[2].running_mean, twin.nn[2].running_mean, atol=1e-5)
    assert torch.allclose(conv.nn[2].running_var, twin.nn[2].running_var, rtol=1e-4, atol=1e-6)
    blk = sparse.ResDynBlock(16, kernel_size=4, dilation=2, norm_type='batch', epsilon=0.0).cuda()
    y, b = blk(x, batch)
    assert y.shape == x.shape and b is batch

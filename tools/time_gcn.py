"""Time network.DenseDeepGCN at train_rnr.py's default size on the GPU (V = 7500, 20 blocks, k = 16, stochastic dilation) --
the `v_feature = gcn(gcn_input)` of train_rnr.py:490 -- and its pieces (kNN distance GEMM + top-k, EdgeConv)."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from relightable_nr_b200.dropin import network
from relightable_nr_b200.dropin.gcn_lib.dense import torch_edge
import make_scene


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    opt = types.SimpleNamespace(n_filters=64, kernel_size=16, act_type='relu', norm_type='batch', bias=True, epsilon=0.2, stochastic=True,
                                conv_type='edge', n_blocks=20, num_v_gcn=7500, out_channels_gcn=512, in_channels=6, block_type='res')
    torch.manual_seed(0)
    gcn = network.DenseDeepGCN(opt).cuda().train()
    v = torch.tensor(make_scene.grid_sphere(75, 100)[0], dtype=torch.float32).cuda()
    inp = types.SimpleNamespace(pos=v, x=v)
    with torch.no_grad():
        t = timeit(lambda: gcn(inp))
    print('DenseDeepGCN forward (V=7500, 20 blocks, no_grad): %.2f ms' % t)
    t = timeit(lambda: gcn(inp))
    print('DenseDeepGCN forward (autograd graph recorded, as train_rnr.py:490 runs it): %.2f ms' % t)
    x = torch.randn(1, 7500, 64, 1, device='cuda')
    for d in (1, 10, 19):
        k = 16 * d
        t_all = timeit(lambda: torch_edge.dense_knn_matrix(x, k))
        t_dist = timeit(lambda: torch_edge.pairwise_distance(x.squeeze(-1)))
        dist = -torch_edge.pairwise_distance(x.squeeze(-1))
        t_topk = timeit(lambda: torch.topk(dist, k=k))
        print('kNN graph, dilation %2d (top-%3d of 7500): total %.3f ms = distance %.3f + top-k %.3f' % (d, k, t_all, t_dist, t_topk))


if __name__ == '__main__':
    main()

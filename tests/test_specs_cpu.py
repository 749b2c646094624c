"""Host-side accounting of the U-Net (no GPU): the live-layer table the engine executes and the FLOP counts bench.py's roofline
uses must match SURVEY.md Appendix A / section 8d (22 live layers; RNR 428.7 GFLOP forward per 512^2 view, DNR-16 553.6)."""
import pytest

torch = pytest.importorskip('torch')


def _specs(in_ch, out_ch, nf0, H=512):
    from relightable_nr_b200.engine.unet import unet_layer_specs
    return unet_layer_specs(in_ch, out_ch, nf0, 5, 8 * nf0, H, H)


def _gflop(specs):
    from relightable_nr_b200.engine.unet import UNetEngine
    return sum(UNetEngine.layer_flops(sp, 1) for sp in specs) / 1e9


def test_rnr_layer_table_and_flops():
    specs = _specs(108, 78, 64)
    assert len(specs) == 22
    kinds = [sp.kind for sp in specs]
    assert kinds.count('ct') == 5 and kinds.count('c4s2') == 5 and kinds.count('c3') == 12
    assert specs[0].name == 'in' and specs[-1].name == 'out' and specs[-1].src == ['x0', 'y0']
    assert [sp.cout for sp in specs[:11]] == [64, 64, 128, 128, 256, 256, 512, 512, 512, 512, 512]
    # innermost block has no BatchNorm but biases; the last layer has a bias and no activation
    inner = [sp for sp in specs if sp.name.startswith('b4.')]
    assert all(sp.bn_key is None and sp.b_key is not None for sp in inner)
    assert specs[-1].bn_key is None and specs[-1].slope is None
    assert abs(_gflop(specs) - 428.7) < 0.1
    # fwd + dgrad (first layer: 24 of 108 input channels) + wgrad
    f = _gflop(specs)
    first = 2.0 * 512 * 512 * 108 * 64 * 9 / 1e9
    assert abs((3 * f - first * 84 / 108) - 1260.7) < 0.5


def test_dnr_flops():
    assert abs(_gflop(_specs(16, 3, 80)) - 553.6) < 0.1


def test_quarter_resolution_scales_flops_by_four():
    assert abs(_gflop(_specs(108, 78, 64, 256)) * 4 - _gflop(_specs(108, 78, 64))) < 1e-6


def test_gcn_lib_sparse_names_and_scatter_path():
    """gcn_lib.sparse exposes the reference's class / function names (gcn_lib/sparse/__init__.py star imports) and its edge-list
    EdgConv equals a naive per-node loop on CPU tensors (the scatter path; the fused CUDA path is checked in tests/test_gcn_gpu.py)."""
    import torch
    from relightable_nr_b200.dropin.gcn_lib import sparse
    for name in ('MRConv', 'EdgConv', 'GraphConv', 'DynConv', 'ResDynBlock', 'DenseDynBlock', 'Dilated', 'DilatedKnnGraph', 'pairwise_distance',
                 'knn_matrix', 'knn_graph_matrix', 'act_layer', 'norm_layer', 'MultiSeq', 'MLP'):
        assert hasattr(sparse, name), name
    torch.manual_seed(0)
    conv = sparse.EdgConv(5, 7, 'relu', None, True)
    x = torch.randn(12, 5)
    ei = sparse.knn_graph_matrix(x, 3, torch.zeros(12, dtype=torch.long))
    out = conv(x, ei)
    for i in range(12):
        js = ei[0][ei[1] == i]
        want = torch.stack([conv.nn(torch.cat([x[i], x[j] - x[i]])[None])[0] for j in js]).max(0)[0]
        assert torch.allclose(out[i], want, atol=1e-6)


def test_packed_view_cache_round_trip(tmp_path):
    """relightable_nr_b200.precompute.PackedViewCache: every map of every view survives the one-file cache bit for bit (host side;
    the pinned / device staging is exercised in tests/test_scripts_gpu.py)."""
    import numpy as np
    import torch
    from relightable_nr_b200.precompute import PackedViewCache
    rng = np.random.RandomState(0)
    views = [{'uv_map': torch.from_numpy(rng.rand(6, 5, 2).astype(np.float32)), 'alpha_map': torch.from_numpy((rng.rand(6, 5) > 0.5).astype(np.float32)),
              'TBN_map': torch.from_numpy(rng.randn(6, 5, 3, 3).astype(np.float32)), 'idx': np.array(i, dtype=np.int64)} for i in range(3)]
    cache = PackedViewCache.write(str(tmp_path / 'views.rnrcache'), views, names=['a', 'b', 'c'])
    again = PackedViewCache(str(tmp_path / 'views.rnrcache'))
    assert len(again) == 3 and again.header['names'] == ['a', 'b', 'c']
    for i, v in enumerate(views):
        got = again.view_numpy(i)
        assert set(got) == set(v)
        for k in v:
            want = v[k].numpy() if isinstance(v[k], torch.Tensor) else v[k]
            assert got[k].dtype == want.dtype and np.array_equal(got[k], want), k
    with open(str(tmp_path / 'junk'), 'wb') as fh:
        fh.write(b'not a cache at all')
    import pytest
    with pytest.raises(ValueError):
        PackedViewCache(str(tmp_path / 'junk'))

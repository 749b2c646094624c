"""tcgen05 weight-gradient kernels (wgrad_halo.cu: halo reuse + one TMEM accumulator per tap; wgrad_tc.cu: one box pair per tap,
plain and swapped operands; GEMM-order scratch + un-transpose) against the SIMT validation kernel on identical fp16 / bf16 operands:
two engines share parameters and input, differ only in the weight-gradient implementation.  Tolerance: relative L2 <= 2e-3 per
weight tensor (fp32 accumulation in a different order; the operands are bit-identical)."""
import os

import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _run(wgrad_impl, H, nf0, in_ch, out_ch, env=None):
    from oracle.unet import make_unet_state_dict
    from relightable_nr_b200.engine.unet import UNetEngine, unet_layer_specs
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        sd = make_unet_state_dict(in_ch, out_ch, nf0, num_down=5, seed=0)
        dev = torch.device('cuda:0')
        params = {k: v.to(dev).contiguous() for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
        buffers = {k: v.to(dev).clone() for k, v in sd.items() if 'running' in k}
        specs = unet_layer_specs(in_ch, out_ch, nf0, 5, 8 * nf0, H, H)
        eng = UNetEngine(specs, params, buffers, 1, in_ch, dev, impl='tc', input_grad_range=(in_ch - 24, in_ch), wgrad_impl=wgrad_impl)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, in_ch, H, H, generator=g).cuda()
    R = (torch.randn(1, out_ch, H, H, generator=g) / (H * H)).cuda()
    eng.set_input_nchw(x)
    eng.forward(training=True)
    eng.backward_from_nchw(R)
    torch.cuda.synchronize()
    return {sp.name: eng.grad_view(sp.w_key).clone() for sp in specs}


@pytest.mark.parametrize('env', [None, {'RNR_WGRAD_HALO': '2'}, {'RNR_WGRAD_HALO': '0'}], ids=['default', 'halo-everywhere', 'per-tap'])
def test_tc_weight_gradients_match_simt(env):
    H, nf0, in_ch, out_ch = 256, 64, 108, 78
    ref = _run('simt', H, nf0, in_ch, out_ch)
    got = _run('tc', H, nf0, in_ch, out_ch, env)
    worst, wn = 0.0, None
    for name in ref:
        e = rel_l2(got[name], ref[name])
        if e > worst:
            worst, wn = e, name
    print('worst weight-gradient relative L2 vs SIMT: %.3e (%s)' % (worst, wn))
    assert worst <= 2e-3, (worst, wn)

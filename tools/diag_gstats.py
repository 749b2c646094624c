"""Per-layer dgamma / dbeta of the U-Net engine with the BatchNorm-backward sums fused into the data-gradient epilogue vs the
separate reduction pass, in backward order (GPU box).  usage: python tools/diag_gstats.py [nf0] [H] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.unet import make_unet_state_dict
from relightable_nr_b200.engine.unet import UNetEngine, unet_layer_specs

nf0 = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1
in_ch, out_ch = 108, 78
sd = make_unet_state_dict(in_ch, out_ch, nf0, num_down=5, seed=0)
g = torch.Generator().manual_seed(1)
x = torch.randn(N, in_ch, H, H, generator=g).cuda()
R = (torch.randn(N, out_ch, H, H, generator=g) / (H * H)).cuda()
dev = torch.device('cuda:0')
res = {}
for mode in ('0', '1'):
    os.environ['RNR_BN_BWD_FUSED'] = mode
    params = {k: v.to(dev).contiguous() for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
    buffers = {k: v.to(dev).clone() for k, v in sd.items() if 'running' in k}
    specs = unet_layer_specs(in_ch, out_ch, nf0, 5, 8 * nf0, H, H)
    eng = UNetEngine(specs, params, buffers, N, in_ch, dev, impl='tc', input_grad_range=(84, 108), wgrad_impl='tc')
    eng.set_input_nchw(x)
    eng.forward(training=True, drop_masks=None)
    eng.backward_from_nchw(R)
    torch.cuda.synchronize()
    res[mode] = ({k: eng.grad_view(k).clone() for k in eng.grad_slices}, set(eng.gstat_layers),
                 {sp.name: eng.gz[sp.name].t.float().clone() for sp in eng.specs})
fused = res['1'][1]
for sp in reversed(specs):
    if sp.bn_key is None:
        continue
    a_g, b_g = res['0'][0][sp.bn_key + '.weight'], res['1'][0][sp.bn_key + '.weight']
    a_b, b_b = res['0'][0][sp.bn_key + '.bias'], res['1'][0][sp.bn_key + '.bias']
    za, zb = res['0'][2][sp.name], res['1'][2][sp.name]
    rl = lambda u, v: ((u - v).norm() / v.norm().clamp_min(1e-30)).item()
    print('%-10s %s  dgamma rel %.2e  dbeta rel %.2e  gz rel %.2e   |dbeta| %.3e' % (
        sp.name, 'FUSED' if sp.name in fused else '     ', rl(b_g, a_g), rl(b_b, a_b), rl(zb, za), a_b.norm().item()))

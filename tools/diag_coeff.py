"""Diagnostic: SH-coefficient gradient of the RNR step, split by loss term, product vs oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pixel_ops as P  # noqa: E402
from oracle.rnr_step import rnr_forward, state_from_pipeline  # noqa: E402
from relightable_nr_b200.pipeline import RNRPipeline, synthetic_view  # noqa: E402
from tests.util import cosine, rel_l2  # noqa: E402

pipe = RNRPipeline(device='cuda:0', img_size=64, texture_size=64, texture_num_ch=24, mipmap_level=3, nf0=16, sh_lmax=4,
                   num_l_samples=512, lp_recon_h=16, lp_recon_w=32, dropout=False)
view = synthetic_view(64, view_idx=5, device='cuda:0')
state = state_from_pipeline(pipe)
final, rays_lt, alpha = pipe.forward(view)
loss, parts = pipe.losses(view, final, rays_lt, alpha)
ours = {}
for k in ('lighting', 'rn', 'chrom'):
    pipe.lighting_model.coeff.grad = None
    parts[k].backward(retain_graph=True)
    g = pipe.lighting_model.coeff.grad
    ours[k] = None if g is None else g[0].detach().cpu().clone()

vc = {k: v.detach().cpu() for k, v in view.items()}
coeff = state['coeff'].clone().requires_grad_(True)
f2, lt2, a2 = rnr_forward(state['textures'], state['unet_sd'], coeff, state['basis_val_recon'], state['lp_hw'], state['pivots_s'],
                          state['pivots_d'], vc)
l_est = P.reconstruct_sh(coeff, state['basis_val'])
m = state['l_mask']
l_init = state['l_init']
w = state['w']
ref_parts = {}
ref_parts['lighting'] = (l_init[m] - l_est[m]).abs().sum() / m.float().sum() * w['lighting'] + \
                        (l_init[~m] - l_est[~m]).abs().sum() / (~m).float().sum() * w['lighting_uncovered']
a = a2[:, :, 5:-5, 5:-5]
ref_parts['rn'] = torch.nn.functional.l1_loss((f2[:, :, 5:-5, 5:-5] * a).reshape(-1), (vc['img_gt'][:, :, 5:-5, 5:-5] * a).reshape(-1))
for k in ('lighting', 'rn'):
    coeff.grad = None
    ref_parts[k].backward(retain_graph=True)
    r = coeff.grad.clone()
    o = ours[k]
    print(k, 'loss ours %.6f ref %.6f' % (parts[k].item(), ref_parts[k].item()),
          'cos %.5f rel %.3e |ours| %.3e |ref| %.3e' % (cosine(o, r), rel_l2(o, r), o.norm().item(), r.norm().item()))
    print('  ours[:3]', o[:3].flatten().tolist())
    print('  ref [:3]', r[:3].flatten().tolist())
print('chrom grad on coeff (should be None):', ours['chrom'])

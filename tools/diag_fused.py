"""Per-tensor comparison of the fused step's gradients with the module path (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightable_nr_b200.pipeline import RNRPipeline, synthetic_view
from tests.util import cosine

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = dict(device='cuda:0', img_size=size, texture_size=64, texture_num_ch=24, mipmap_level=3, nf0=16, sh_lmax=4,
           num_l_samples=512, lp_recon_h=16, lp_recon_w=32, dropout=False)
pipe = RNRPipeline(**cfg)
view = synthetic_view(size, view_idx=5, device='cuda:0')
final, rays_lt, alpha = pipe.forward(view)
loss_m, _ = pipe.losses(view, final, rays_lt, alpha)
loss_m.backward()
gm = {k: p.grad.clone() for k, p in pipe.render_net.named_parameters() if p.grad is not None}
pipe.optimizer.zero_grad(set_to_none=True)
loss_f, final_f = pipe.fused.train_step(view, step_optimizer=False)
torch.cuda.synchronize()
print('loss', loss_m.item(), loss_f.item())
for k, p in pipe.render_net.named_parameters():
    if k in gm:
        c = cosine(p.grad, gm[k])
        if c < 0.9999:
            print('%-60s cos %.5f  |module| %.3e |fused| %.3e' % (k, c, gm[k].norm().item(), p.grad.norm().item()))

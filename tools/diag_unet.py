"""Per-layer diagnostics of the U-Net engine against the CPU oracle (run on the GPU box).

usage: python tools/diag_unet.py [simt|tc] [nf0] [H] [N] [in_ch] [out_ch]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.unet import make_unet_state_dict, unet_forward
from relightable_nr_b200.engine.unet import UNetEngine, unet_layer_specs


def main():
    impl = sys.argv[1] if len(sys.argv) > 1 else 'simt'
    nf0 = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    N = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    in_ch = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    out_ch = int(sys.argv[6]) if len(sys.argv) > 6 else 6
    wimpl = sys.argv[7] if len(sys.argv) > 7 else 'simt'
    q16 = (sys.argv[8] == 'q16') if len(sys.argv) > 8 else True
    num_down = 5
    sd = make_unet_state_dict(in_ch, out_ch, nf0, num_down=num_down, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, in_ch, H, H, generator=g)
    R = torch.randn(N, out_ch, H, H, generator=g) / (H * H)

    dev = torch.device('cuda:0')
    params = {k: v.to(dev).contiguous() for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
    buffers = {k: v.to(dev).clone() for k, v in sd.items() if 'running' in k}
    specs = unet_layer_specs(in_ch, out_ch, nf0, num_down, 8 * nf0, H, H)
    adt = {'fp16': 0, 'bf16': 1}[os.environ.get('RNR_ACT_DTYPE', 'fp16')]
    eng = UNetEngine(specs, params, buffers, N, in_ch, dev, impl=impl, input_grad_range=(0, in_ch), wgrad_impl=wimpl, act_dtype=adt)
    eng.set_input_nchw(x.to(dev))
    eng.forward(training=True)
    torch.cuda.synchronize()
    gates = {k: v.cpu() for k, v in eng.gate_masks().items()} if q16 else None

    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k else v) for k, v in sd.items()}
    xg = x.clone().requires_grad_(True)
    t0 = time.time()
    pre, acts = unet_forward(sdg, xg, num_down=num_down, return_acts=True, q16=q16, gates=gates)
    for a in acts.values():
        a.retain_grad()
    ref = torch.tanh(pre)
    (ref * R).sum().backward()
    print('oracle fwd+bwd %.2fs' % (time.time() - t0))
    print('--- forward activations (engine vs oracle): max-abs / rel-l2')
    for sp in specs[:-1]:
        a = eng.acts[sp.dst].interior_tensor().float().permute(0, 3, 1, 2).cpu()
        r = acts[sp.dst].detach()
        d = (a - r).abs().max().item()
        rl = ((a - r).norm() / (r.norm() + 1e-30)).item()
        # halo check: reflect
        full = eng.acts[sp.dst].t.float().permute(0, 3, 1, 2).cpu()
        rp = torch.nn.functional.pad(a, (1, 1, 1, 1), mode='reflect')
        hd = (full - rp).abs().max().item()
        print('%-10s %-5s C=%4d HxW=%3d  max-abs %.3e  rel-l2 %.3e  halo-err %.1e' % (sp.name, sp.kind, sp.cout, sp.Ho, d, rl, hd))
    out = eng.output_nchw().cpu()
    print('out: max-abs %.3e' % (out - ref.detach()).abs().max().item())

    gx = eng.backward_from_nchw(R.to(dev))
    torch.cuda.synchronize()
    print('--- backward: gz (grad wrt raw) is internal; compare param grads + grad wrt activations')
    for sp in reversed(specs):
        for key in (sp.w_key, sp.b_key, (sp.bn_key + '.weight') if sp.bn_key else None, (sp.bn_key + '.bias') if sp.bn_key else None):
            if key is None:
                continue
            gm = eng.grad_view(key).cpu()
            gr = sdg[key].grad
            rl = ((gm - gr).norm() / (gr.norm() + 1e-30)).item()
            cs = (torch.dot(gm.flatten(), gr.flatten()) / (gm.norm() * gr.norm() + 1e-30)).item()
            print('%-10s %-40s rel-l2 %.3e cos %.6f |ref| %.3e' % (sp.name, key, rl, cs, gr.norm().item()))
    rl = ((gx.cpu() - xg.grad).norm() / (xg.grad.norm() + 1e-30)).item()
    print('input grad rel-l2 %.3e' % rl)
    print('gpu launches', eng.gpu_launches)

    # timing
    for _ in range(2):
        eng.forward(training=True)
        eng._backward_layers()
    torch.cuda.synchronize()
    e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    eng.forward(training=True)
    e1.record()
    eng._backward_layers()
    e2.record()
    torch.cuda.synchronize()
    print('time fwd %.3f ms  bwd %.3f ms' % (e0.elapsed_time(e1), e1.elapsed_time(e2)))


if __name__ == '__main__':
    main()

"""Per-layer timing of the U-Net engine at a given size (run on the GPU box).

usage: python tools/perf_unet.py [impl] [nf0] [H] [N] [in_ch] [out_ch] [wgrad_impl]
Prints, per layer, the forward conv time and achieved TFLOP/s (2*MAC, live FLOPs), then totals.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.unet import make_unet_state_dict
from relightable_nr_b200 import _lib
from relightable_nr_b200.engine.unet import UNetEngine, unet_layer_specs


def layer_flops(sp, N):
    k = 3 if sp.kind == 'c3' else 4
    cin = sum(sp.cin)
    if sp.kind == 'ct':
        return 2.0 * N * sp.H * sp.W * cin * sp.cout * 16
    return 2.0 * N * sp.Ho * sp.Wo * cin * sp.cout * k * k


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    impl = sys.argv[1] if len(sys.argv) > 1 else 'tc'
    nf0 = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    N = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    in_ch = int(sys.argv[5]) if len(sys.argv) > 5 else 108
    out_ch = int(sys.argv[6]) if len(sys.argv) > 6 else 78
    wimpl = sys.argv[7] if len(sys.argv) > 7 else 'simt'
    sd = make_unet_state_dict(in_ch, out_ch, nf0, num_down=5, seed=0)
    dev = torch.device('cuda:0')
    params = {k: v.to(dev).contiguous() for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
    buffers = {k: v.to(dev).clone() for k, v in sd.items() if 'running' in k}
    specs = unet_layer_specs(in_ch, out_ch, nf0, 5, 8 * nf0, H, H)
    gr = (in_ch - 24, in_ch) if in_ch > 24 else (0, in_ch)
    eng = UNetEngine(specs, params, buffers, N, in_ch, dev, impl=impl, input_grad_range=gr, wgrad_impl=wimpl)
    x = torch.randn(N, in_ch, H, H, device=dev)
    eng.set_input_nchw(x)
    eng.forward(training=True)
    R = torch.randn(N, out_ch, H, H, device=dev) / (H * H)
    eng.backward_from_nchw(R)
    torch.cuda.synchronize()
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    tot_f = tot_d = tot_w = 0.0
    tot_flop = 0.0
    print('%-10s %-5s %5s->%-4s %4s | fwd ms  TF/s | dgrad ms TF/s | wgrad ms TF/s' % ('layer', 'kind', 'cin', 'cout', 'Ho'))
    only = os.environ.get('PERF_LAYERS')
    for sp in specs:
        if only and sp.name not in only.split(','):
            continue
        st = eng.layers[sp.name]
        fl = layer_flops(sp, N)
        tf = timeit(lambda: [L.rnr_conv_run(p.h, s) for p in st.fwd_plans])
        td = timeit(lambda: [L.rnr_conv_run(p.h, s) for p in st.dgrad_plans]) if st.dgrad_plans else 0.0
        tw = timeit(lambda: L.rnr_wgrad_run(st.wgrad_plan.h, s), iters=2)
        tot_f += tf; tot_d += td; tot_w += tw; tot_flop += fl
        dfl = fl if sp.name != 'in' else fl * (gr[1] - gr[0]) / in_ch
        print('%-10s %-5s %5d->%-4d %4d | %6.3f %6.1f | %6.3f %6.1f | %7.3f %6.1f' % (
            sp.name, sp.kind, sum(sp.cin), sp.cout, sp.Ho, tf, fl / tf / 1e9, td, (dfl / td / 1e9) if td else 0.0, tw, fl / tw / 1e9))
    print('conv totals: fwd %.3f ms (%.1f TF/s)  dgrad %.3f ms  wgrad %.3f ms   fwd GFLOP %.1f' % (
        tot_f, tot_flop / tot_f / 1e9, tot_d, tot_w, tot_flop / 1e9))
    t_fwd = timeit(lambda: eng.forward(training=True), iters=5)
    t_bwd = timeit(lambda: eng._backward_layers(), iters=2)
    t_wp = timeit(lambda: eng.prepare_weights(backward=True), iters=5)
    print('engine forward (incl. weight prep, BN passes) %.3f ms; backward %.3f ms; weight prep alone %.3f ms' % (t_fwd, t_bwd, t_wp))


if __name__ == '__main__':
    main()

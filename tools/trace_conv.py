"""Timeline (clock64 stamps) of the halo conv kernel's warp roles for one layer.  usage: python tools/trace_conv.py <layer> [fwd|dgrad]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.unet import make_unet_state_dict
from relightable_nr_b200 import _lib
from relightable_nr_b200.engine.unet import UNetEngine, unet_layer_specs

layer = sys.argv[1] if len(sys.argv) > 1 else 'b1.down1'
kind = sys.argv[2] if len(sys.argv) > 2 else 'fwd'
H, nf0, in_ch, out_ch = 512, 64, 108, 78
sd = make_unet_state_dict(in_ch, out_ch, nf0, num_down=5, seed=0)
dev = torch.device('cuda:0')
params = {k: v.to(dev).contiguous() for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
buffers = {k: v.to(dev).clone() for k, v in sd.items() if 'running' in k}
specs = unet_layer_specs(in_ch, out_ch, nf0, 5, 8 * nf0, H, H)
eng = UNetEngine(specs, params, buffers, 1, in_ch, dev, impl='tc', input_grad_range=(84, 108), wgrad_impl='tc')
eng.set_input_nchw(torch.randn(1, in_ch, H, H, device=dev))
eng.forward(training=True)
eng.backward_from_nchw(torch.randn(1, out_ch, H, H, device=dev) / (H * H))
torch.cuda.synchronize()
L = _lib.lib()
s = torch.cuda.current_stream().cuda_stream
st = eng.layers[layer]
plan = (st.fwd_plans if kind == 'fwd' else st.dgrad_plans)[0]
for _ in range(3):
    L.rnr_conv_run(plan.h, s)
torch.cuda.synchronize()
buf = torch.zeros(4 * 4 * 64, dtype=torch.int64, device=dev)
_lib.check(L.rnr_debug_set_trace(buf.data_ptr()))
L.rnr_conv_run(plan.h, s)
torch.cuda.synchronize()
_lib.check(L.rnr_debug_set_trace(None))
t = buf.cpu().view(4, 4, 64)
for cta in range(2):
    t0 = int(t[cta, 3, 0])
    print('CTA %d (cycles since kernel start)' % cta)
    for role, name in ((3, 'kernel start, then tile-0 epilogue per slab [ld done, staged, stored, stats done]'), (0, 'producer tile[start,end]'), (1, 'mma tile[acc free, first A, issued]'), (2, 'epilogue tile[acc full, done]')):
        v = [int(x) - t0 for x in t[cta, role] if int(x) != 0]
        print('  %-38s %s' % (name, v))

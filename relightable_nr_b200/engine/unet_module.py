"""Autograd bridge between the drop-in ``Unet`` parameter tree and the tcgen05 engine.

``UnetRunner(module)(x, apply_tanh)`` is what ``Unet.forward`` / ``RenderingNet.forward`` call
(network.py:251-253; pytorch_prototyping/pytorch_prototyping.py:532-536).  It keeps one UNetEngine
per (device, N, H, W, needs-grad, tanh) -- plans, TMA descriptors and all activation storage are
created once and reused every step -- and exposes the run as a ``torch.autograd.Function`` so the
caller's ``loss.backward()`` / ``torch.optim`` work unchanged.
"""
import weakref

import torch

from .unet import UNetEngine, unet_layer_specs


def _live_params(unet):
    """{engine key: Parameter} for the live layers, {key: buffer} for BN running stats."""
    sd_params = dict(unet.named_parameters(remove_duplicate=False))
    sd_bufs = dict(unet.named_buffers())
    return sd_params, sd_bufs


class _UnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, eng, training_bn, drop_masks, x, *params):
        eng.set_input_nchw(x)
        eng.forward(training=training_bn, drop_masks=drop_masks, need_backward_prep=eng.need_backward)
        out = eng.output_nchw()
        eng.version += 1
        ctx.eng = eng
        ctx.version = eng.version
        ctx.runner = runner
        ctx.n_params = len(params)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.eng
        if eng.version != ctx.version:
            raise RuntimeError('U-Net engine buffers were overwritten by a later forward of the same shape before backward(); '
                               'run backward before the next grad-enabled forward')
        gx = eng.backward_from_nchw(grad_out.contiguous().float())
        flat = eng.grad_flat.clone()          # hand out a private copy: the engine buffer is reused next step
        grads = []
        for key in ctx.runner.param_keys:
            if key in eng.grad_slices:
                o, n = eng.grad_slices[key]
                grads.append(flat[o:o + n].view(eng.params[key].shape))
            else:
                grads.append(None)
        gin = None
        if gx is not None:
            r0, r1 = eng.input_grad_range
            if (r0, r1) == (0, eng.in_channels):
                gin = gx
            else:
                gin = torch.zeros((eng.N, eng.in_channels, eng.H, eng.W), dtype=torch.float32, device=gx.device)
                gin[:, r0:r1] = gx
        return (None, None, None, None, gin, *grads)


class UnetRunner:
    def __init__(self, unet):
        self._unet = weakref.ref(unet)
        self._engines = {}
        self.param_keys = None
        #: channels of the input that need a gradient (None = all).  network.RenderingNet narrows this to the
        #: neural-texture channels, the only differentiable part of the 108-channel RNR input.
        self.input_grad_range = None
        #: how the Dropout2d channel masks are drawn.  'fast': one uniform draw for all 21 live layers.  'reference': one
        #: bernoulli_ per Dropout2d call of the reference's forward, IN ITS ORDER -- including the 22 calls of the dead GCN pass of
        #: the outermost block (pytorch_prototyping.py:407-415) -- from torch's CUDA generator, so that under
        #: the same torch.manual_seed the live layers get bit-identical masks to the reference module running on the same GPU.
        self.rng_order = 'fast'

    def _engine(self, x, need_backward, apply_tanh):
        unet = self._unet()
        cfg = unet._cfg
        N, Cin, H, W = x.shape
        if Cin != cfg['in_channels']:
            raise ValueError('expected %d input channels, got %d' % (cfg['in_channels'], Cin))
        rng = self.input_grad_range if (need_backward and x.requires_grad) else None
        if need_backward and x.requires_grad and rng is None:
            rng = (0, Cin)
        return self.engine_for(x.device, N, H, W, need_backward, apply_tanh, rng)

    def engine_for(self, device, N, H, W, need_backward, apply_tanh, rng):
        """The (cached) engine for an input shape.  Also the entry point of the fused step (relightable_nr_b200/fused.py),
        whose producer kernel writes the first convolution's operand in place instead of handing over an NCHW tensor."""
        unet = self._unet()
        cfg = unet._cfg
        Cin = cfg['in_channels']
        device = torch.device(device)
        key = (device, N, H, W, need_backward, apply_tanh, rng)
        params, bufs = _live_params(unet)
        eng = self._engines.get(key)
        if eng is not None:
            # parameters may have been re-created (module.to(), load_state_dict keeps them): re-bind by identity
            if all(eng.params[k] is params[k] for k in eng.params):
                return eng
            del self._engines[key]
        if H % (2 ** cfg['num_down']) or W % (2 ** cfg['num_down']) or min(H, W) < 2 ** (cfg['num_down'] + 1):
            raise ValueError('input size %dx%d must be a multiple of %d and at least %d' % (
                H, W, 2 ** cfg['num_down'], 2 ** (cfg['num_down'] + 1)))
        specs = unet_layer_specs(cfg['in_channels'], cfg['out_channels'], cfg['nf0'], cfg['num_down'], cfg['max_channels'], H, W)
        live = {}
        for sp in specs:
            for k in (sp.w_key, sp.b_key, sp.bn_key + '.weight' if sp.bn_key else None, sp.bn_key + '.bias' if sp.bn_key else None):
                if k is not None:
                    live[k] = params[k]
        for k, p in live.items():
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise TypeError('U-Net parameter %s must be a contiguous fp32 CUDA tensor (librnr_b200 has no CPU path)' % k)
        live_bufs = {k: v for k, v in bufs.items() if 'running' in k or k.endswith('num_batches_tracked')}
        eng = UNetEngine(specs, live, live_bufs, N, Cin, device, impl='tc', input_grad_range=rng,
                         need_backward=need_backward, final_tanh=apply_tanh)
        eng.version = 0
        # bound memory: keep at most 4 shapes alive
        if len(self._engines) >= 4:
            self._engines.pop(next(iter(self._engines)))
        self._engines[key] = eng
        return eng

    def __call__(self, x, apply_tanh):
        unet = self._unet()
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise TypeError('Unet input must be a CUDA tensor (librnr_b200 has no CPU path)')
        x = x.float().contiguous()
        params, bufs = _live_params(unet)
        if self.param_keys is None:
            cfg = unet._cfg
            probe = unet_layer_specs(cfg['in_channels'], cfg['out_channels'], cfg['nf0'], cfg['num_down'], cfg['max_channels'], 64, 64)
            keys = []
            for sp in probe:
                for k in (sp.w_key, sp.b_key, sp.bn_key + '.weight' if sp.bn_key else None, sp.bn_key + '.bias' if sp.bn_key else None):
                    if k is not None:
                        keys.append(k)
            self.param_keys = keys
        plist = [params[k] for k in self.param_keys]
        need_backward = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in plist))
        eng = self._engine(x, need_backward, apply_tanh)
        training_bn, drop_masks = self.step_state(eng)
        return _UnetFn.apply(self, eng, training_bn, drop_masks, x, *plist)

    def step_state(self, eng):
        """Per-forward host work of the nn.Module semantics: (BatchNorm in training mode?, Dropout2d channel masks or None);
        bumps ``num_batches_tracked`` like nn.BatchNorm2d.forward does."""
        unet = self._unet()
        bn_mod = unet.in_layer[1]
        drop_mod = unet.in_layer[3]
        training_bn = bool(bn_mod.training)
        # the engine applies ONE BatchNorm mode and ONE Dropout2d (mode, p) to every live layer: refuse mixed settings rather
        # than silently ignoring a per-module .eval() / a different p (the dead `fuse` branch is never executed)
        for name, m in unet.named_modules():
            if '.fuse' in name:
                continue
            if isinstance(m, torch.nn.BatchNorm2d) and bool(m.training) != training_bn:
                raise NotImplementedError('Unet: BatchNorm2d modules are in mixed train()/eval() modes (%s)' % name)
            if isinstance(m, torch.nn.Dropout2d) and (bool(m.training) != bool(drop_mod.training) or float(m.p) != float(drop_mod.p)):
                raise NotImplementedError('Unet: Dropout2d modules differ in mode or p (%s)' % name)
        drop_masks = None
        if drop_mod.training and drop_mod.p > 0 and self.rng_order == 'reference':
            p = float(drop_mod.p)

            def draw(c):
                # torch's feature_dropout: noise = empty([N, C, 1, 1]).bernoulli_(1 - p).div_(1 - p)
                return torch.empty((eng.N, c, 1, 1), dtype=torch.float32, device=eng.device).bernoulli_(1 - p).div_(1 - p).view(eng.N, c).contiguous()

            cout = {sp.name: sp.cout for sp in eng.specs}
            nd = unet._cfg['num_down']
            body = ['b0.down1', 'b0.down2']
            inner = []
            for i in range(1, nd):
                inner += ['b%d.down1' % i, 'b%d.down2' % i]
            for i in reversed(range(1, nd)):
                inner += ['b%d.up1' % i, 'b%d.up2' % i]
            tail = ['b0.up1', 'b0.up2']
            drop_masks = {'in': draw(cout['in'])}
            if unet.use_gcn:
                # the dead pass of the outermost block: down (2 draws), fuse (2: its DownBlock keeps inner_nc + out_channels_gcn
                # channels in the middle), the whole submodule, up (2) -- drawn and discarded, only to advance the generator
                fuse_mid, fuse_out = unet.unet_block.fuse.net[1].out_channels, unet.unet_block.fuse.net[6].out_channels
                for c in [cout[n] for n in body] + [fuse_mid, fuse_out] + [cout[n] for n in inner] + [cout[n] for n in tail]:
                    draw(c)
            for n in body + inner + tail:
                drop_masks[n] = draw(cout[n])
        elif drop_mod.training and drop_mod.p > 0:
            p = float(drop_mod.p)
            names = [sp.name for sp in eng.specs if sp.drop and sp.dst != 'out']
            chans = [eng.layers[n].spec.cout for n in names]
            r = (torch.rand((eng.N, sum(chans)), device=eng.device) >= p).float() * (1.0 / (1.0 - p))
            drop_masks, o = {}, 0
            for n, c in zip(names, chans):
                drop_masks[n] = r[:, o:o + c].contiguous()
                o += c
        if training_bn:
            # (layers whose statistics are finalized inside the conv kernel count their batches there)
            nbt = [b for k, b in unet.named_buffers() if k.endswith('num_batches_tracked') and '.fuse.' not in k
                   and not eng.counts_batches_in_kernel(k[:-len('.num_batches_tracked')])]
            if nbt:
                torch._foreach_add_(nbt, 1)
        return training_bn, drop_masks

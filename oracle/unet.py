"""TEST INFRASTRUCTURE ONLY -- CPU/plain-PyTorch fp32 restatement of the reference U-Net.

Restates, as a pure function of a reference-format ``state_dict``, what
``network.RenderingNet.forward`` computes (/root/reference/network.py:251-253) through
``pytorch_prototyping.Unet.forward`` (pytorch_prototyping/pytorch_prototyping.py:532-536),
``UnetSkipConnectionBlock.forward`` (:407-429), ``DownBlock`` (:241-274), ``UpBlock`` (:154-199) and
``Conv2dSame`` (:110-121).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this package; the product path never does.

Pinned against the real reference in tests/golden/make_golden.py (run in the build container, where
/root/reference exists) -> tests/golden/unet_*.npz, checked by tests/test_oracle_golden.py.

BatchNorm always uses batch statistics (the reference never evaluates BN with running
stats).  ``drop`` maps a live layer name ('in', 'b0.down1', ...) to a [N, C] multiplicative mask
(0 or 1/(1-p)); None means dropout off (inference mode of test_rnr.py:220-233).
"""
import torch
import torch.nn.functional as F


def _bn(x, sd, key, eps=1e-5):
    if _Q.bn_eval:      # nn.BatchNorm2d.eval(): running statistics (only reached by module.eval() without set_bn_train)
        return F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'], sd[key + '.weight'], sd[key + '.bias'],
                            training=False, eps=eps)
    return F.batch_norm(x, None, None, sd[key + '.weight'], sd[key + '.bias'], training=True, momentum=0.0, eps=eps)


def _drop(x, drop, name):
    if drop is None or drop.get(name) is None:
        return x
    return x * drop[name][:, :, None, None]


def _rpad(x):
    return F.pad(x, (1, 1, 1, 1), mode='reflect')


class _Q:
    """Optional emulation of the engine's storage precision (test-only): weights and post-activation
    tensors rounded to fp16 with a straight-through gradient.  With it the oracle's ReLU/LeakyReLU gates
    agree with the engine's, so backward parity can be checked much tighter than against pure fp32
    (where ~0.1% of gates flip and alone cause a few % relative-L2 gradient difference)."""
    on = False
    bn_eval = False

    gates = None     # optional {layer name: bool mask [N,C,H,W]} of (pre-activation > 0) taken from the engine


def _q(x):
    if not _Q.on:
        return x
    return x + (x.half().float() - x).detach()


def _act(y, slope, name):
    """LeakyReLU(slope) (slope 0 == ReLU).  With _Q.gates the sign decisions come from the engine, so both
    sides differentiate through identical gates (forward values differ only where |y| ~ 1e-3)."""
    if _Q.gates is not None and name in _Q.gates:
        return y * torch.where(_Q.gates[name], torch.ones_like(y), torch.full_like(y, slope))
    return F.leaky_relu(y, slope) if slope != 0.0 else F.relu(y)


def _down(x, sd, pfx, bn, i, drop, acts):
    k1, b1, k2, b2 = ('net.1', 'net.2', 'net.6', 'net.7') if bn else ('net.1', None, 'net.5', None)
    y = F.conv2d(_rpad(x), _q(sd[f'{pfx}.{k1}.weight']), None if bn else sd[f'{pfx}.{k1}.bias'])
    if bn:
        y = _bn(y, sd, f'{pfx}.{b1}')
    y = _q(_drop(_act(y, 0.2, f'b{i}.down1'), drop, f'b{i}.down1'))
    acts[f'd{i}'] = y
    y = F.conv2d(_rpad(y), _q(sd[f'{pfx}.{k2}.weight']), None if bn else sd[f'{pfx}.{k2}.bias'], stride=2)
    if bn:
        y = _bn(y, sd, f'{pfx}.{b2}')
    y = _q(_drop(_act(y, 0.2, f'b{i}.down2'), drop, f'b{i}.down2'))
    acts[f'x{i + 1}'] = y
    return y


def _up(x, sd, pfx, bn, i, drop, acts):
    u1, ub1, u2, ub2 = ('net.0', 'net.1', 'net.4.net.1', 'net.5') if bn else ('net.0', None, 'net.3.net.1', None)
    y = F.conv_transpose2d(x, _q(sd[f'{pfx}.{u1}.weight']), None if bn else sd[f'{pfx}.{u1}.bias'], stride=2, padding=1)
    if bn:
        y = _bn(y, sd, f'{pfx}.{ub1}')
    y = _q(_drop(_act(y, 0.0, f'b{i}.up1'), drop, f'b{i}.up1'))
    acts[f'u{i}'] = y
    y = F.conv2d(_rpad(y), _q(sd[f'{pfx}.{u2}.weight']), None if bn else sd[f'{pfx}.{u2}.bias'])
    if bn:
        y = _bn(y, sd, f'{pfx}.{ub2}')
    y = _q(_drop(_act(y, 0.0, f'b{i}.up2'), drop, f'b{i}.up2'))
    acts[f'y{i}'] = y
    return y


def _block(x, sd, pfx, i, num_down, drop, acts):
    innermost = i == num_down - 1
    bn = not innermost
    y = _down(x, sd, pfx + '.down', bn, i, drop, acts)
    if not innermost:
        y = _block(y, sd, pfx + '.submodule', i + 1, num_down, drop, acts)
    y = _up(y, sd, pfx + '.up', bn, i, drop, acts)
    # highway_mode 'concat' for every block (outermost: network.py:247; inner blocks: default)
    return torch.cat([x, y], 1)


def unet_forward(sd, x, num_down=5, drop=None, prefix='', return_acts=False, q16=False, gates=None, bn_eval=False):
    """sd: state_dict of the reference ``Unet`` (keys relative to ``prefix``).  Returns pre-tanh output.
    q16=True emulates the engine's fp16 storage of weights/activations (see _Q)."""
    _Q.on = bool(q16)
    _Q.gates = gates
    _Q.bn_eval = bool(bn_eval)
    try:
        return _unet_forward(sd, x, num_down, drop, prefix, return_acts)
    finally:
        _Q.on = False
        _Q.gates = None
        _Q.bn_eval = False


def _unet_forward(sd, x, num_down, drop, prefix, return_acts):
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    acts = {}
    y = F.conv2d(_rpad(_q(x)), _q(sd['in_layer.0.net.1.weight']), None)
    y = _bn(y, sd, 'in_layer.1')
    y = _q(_drop(_act(y, 0.2, 'in'), drop, 'in'))
    acts['x0'] = y
    y = _block(y, sd, 'unet_block', 0, num_down, drop, acts)
    y = F.conv2d(_rpad(y), _q(sd['out_layer.0.net.1.weight']), sd['out_layer.0.net.1.bias'])
    if return_acts:
        return y, acts
    return y


def rendering_net_forward(sd, x, num_down=5, drop=None):
    """network.RenderingNet.forward (network.py:251-253): tanh(Unet(x)); sd has the 'net.' prefix."""
    return torch.tanh(unet_forward(sd, x, num_down=num_down, drop=drop, prefix='net.'))


def make_unet_state_dict(in_channels, out_channels, nf0, num_down=5, max_channels=None, seed=0, dtype=torch.float32):
    """Random state_dict with the reference's key names/shapes (used when /root/reference is absent, e.g.
    on the GPU box).  Initialisation mimics PyTorch defaults (kaiming-uniform a=sqrt(5))."""
    g = torch.Generator().manual_seed(seed)
    max_channels = max_channels or 8 * nf0
    sd = {}

    def conv(key, co, ci, k, bias, transpose=False):
        fan_in = (co if transpose else ci) * k * k
        bound = 1.0 / fan_in ** 0.5
        shape = (ci, co, k, k) if transpose else (co, ci, k, k)
        sd[key + '.weight'] = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound
        if bias:
            sd[key + '.bias'] = (torch.rand(co, generator=g, dtype=dtype) * 2 - 1) * bound

    def bn(key, c):
        sd[key + '.weight'] = 1.0 + 0.1 * torch.randn(c, generator=g, dtype=dtype)
        sd[key + '.bias'] = 0.1 * torch.randn(c, generator=g, dtype=dtype)
        sd[key + '.running_mean'] = torch.zeros(c, dtype=dtype)
        sd[key + '.running_var'] = torch.ones(c, dtype=dtype)
        sd[key + '.num_batches_tracked'] = torch.tensor(0)

    conv('in_layer.0.net.1', nf0, in_channels, 3, False)
    bn('in_layer.1', nf0)
    chans = [min(2 ** i * nf0, max_channels) for i in range(num_down)]
    pfx = 'unet_block'
    for i in range(num_down):
        innermost = i == num_down - 1
        outer = chans[i]
        inner = chans[i] if innermost else chans[i + 1]
        has_bn = not innermost
        if has_bn:
            conv(f'{pfx}.down.net.1', outer, outer, 3, False); bn(f'{pfx}.down.net.2', outer)
            conv(f'{pfx}.down.net.6', inner, outer, 4, False); bn(f'{pfx}.down.net.7', inner)
            conv(f'{pfx}.up.net.0', outer, 2 * inner, 4, False, transpose=True); bn(f'{pfx}.up.net.1', outer)
            conv(f'{pfx}.up.net.4.net.1', outer, outer, 3, False); bn(f'{pfx}.up.net.5', outer)
        else:
            conv(f'{pfx}.down.net.1', outer, outer, 3, True)
            conv(f'{pfx}.down.net.5', inner, outer, 4, True)
            conv(f'{pfx}.up.net.0', outer, inner, 4, True, transpose=True)
            conv(f'{pfx}.up.net.3.net.1', outer, outer, 3, True)
        pfx += '.submodule'
    conv('out_layer.0.net.1', out_channels, 2 * nf0, 3, True)
    return sd

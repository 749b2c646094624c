"""Drop-in for ``pytorch_prototyping.pytorch_prototyping`` (2-D U-Net family only).

The classes below keep the reference's constructor signatures, module tree and therefore its
``state_dict`` keys (139 entries for the RNR network, duplicate ``Conv2dSame`` keys included --
pytorch_prototyping/pytorch_prototyping.py:96-121, 124-206, 209-277, 370-429, 432-536), but they
are *parameter holders*: the arithmetic of ``Unet.forward`` is executed by the tcgen05 engine in
``relightable_nr_b200.engine.unet`` on librnr_b200.so.  There is no PyTorch/cuDNN fallback: a CPU
tensor or a missing library raises.

Differences that cannot be observed through outputs or gradients: the dead GCN
branch of the outermost block (``fuse`` + a first evaluation of every inner layer, overwritten at
pytorch_prototyping.py:416-419) is not executed, so ``v_fea`` is accepted and ignored; ``fuse.*``
parameters exist, load and save, and -- as in the reference -- never receive a gradient.
Only ``upsampling_mode='transpose'`` (the only mode the reference scripts instantiate) is built.
"""
import torch
import torch.nn as nn

from ..engine.unet_module import UnetRunner

__all__ = ['Conv2dSame', 'DownBlock', 'UpBlock', 'UnetSkipConnectionBlock', 'Unet']


def _holder_forward(self, *a, **k):
    raise NotImplementedError(
        '%s is a parameter holder of the B200 U-Net engine; call the enclosing Unet / RenderingNet' % type(self).__name__)


class Conv2dSame(nn.Module):
    """ReflectionPad2d + Conv2d holder; re-exports ``weight``/``bias`` like the reference (:117-118)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, padding_layer=nn.ReflectionPad2d):
        super().__init__()
        before = kernel_size // 2
        after = before if kernel_size % 2 else before - 1
        self.net = nn.Sequential(padding_layer((before, after, before, after)),
                                 nn.Conv2d(in_channels, out_channels, kernel_size, bias=bias, stride=1))
        self.weight = self.net[1].weight
        self.bias = self.net[1].bias

    forward = _holder_forward


class DownBlock(nn.Module):
    """pad-conv3x3-[norm]-LeakyReLU-[drop] then pad-conv(4x4 s2 | 3x3 s1)-[norm]-LeakyReLU-[drop] (:209-277)."""

    def __init__(self, in_channels, out_channels, prep_conv=True, middle_channels=None, use_dropout=False,
                 dropout_prob=0.1, norm=nn.BatchNorm2d, stride=2, kernal_size=4):
        super().__init__()
        middle_channels = in_channels if middle_channels is None else middle_channels
        biased = norm is None
        layers = []
        if prep_conv:
            layers += [nn.ReflectionPad2d(1), nn.Conv2d(in_channels, middle_channels, 3, padding=0, stride=1, bias=biased)]
            if norm is not None:
                layers.append(norm(middle_channels, affine=True))
            layers.append(nn.LeakyReLU(0.2, True))
            if use_dropout:
                layers.append(nn.Dropout2d(dropout_prob, False))
        layers += [nn.ReflectionPad2d(1), nn.Conv2d(middle_channels, out_channels, kernal_size, padding=0, stride=stride, bias=biased)]
        if norm is not None:
            layers.append(norm(out_channels, affine=True))
        layers.append(nn.LeakyReLU(0.2, True))
        if use_dropout:
            layers.append(nn.Dropout2d(dropout_prob, False))
        self.net = nn.Sequential(*layers)

    forward = _holder_forward


class UpBlock(nn.Module):
    """ConvTranspose2d(4, s2, p1)-[norm]-ReLU-[drop]-Conv2dSame(3)-[norm]-ReLU-[drop] (:124-206)."""

    def __init__(self, in_channels, out_channels, post_conv=True, use_dropout=False, dropout_prob=0.1,
                 norm=nn.BatchNorm2d, upsampling_mode='transpose'):
        super().__init__()
        if upsampling_mode != 'transpose':
            raise NotImplementedError("only upsampling_mode='transpose' is built (the only mode the reference scripts use)")
        biased = norm is None
        layers = [nn.ConvTranspose2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1, bias=biased)]
        if norm is not None:
            layers.append(norm(out_channels, affine=True))
        layers.append(nn.ReLU(True))
        if use_dropout:
            layers.append(nn.Dropout2d(dropout_prob, False))
        if post_conv:
            layers.append(Conv2dSame(out_channels, out_channels, kernel_size=3, bias=biased))
            if norm is not None:
                layers.append(norm(out_channels, affine=True))
            layers.append(nn.ReLU(True))
            if use_dropout:
                layers.append(nn.Dropout2d(0.1, False))
        self.net = nn.Sequential(*layers)

    forward = _holder_forward


class UnetSkipConnectionBlock(nn.Module):
    def __init__(self, outer_nc, inner_nc, upsampling_mode, norm=nn.BatchNorm2d, submodule=None, use_dropout=False,
                 dropout_prob=0.1, flag_outer=True, gcn=False, out_channels_gcn=512, highway_mode='concat'):
        super().__init__()
        if highway_mode not in ('concat', 'residual', 'no_highway'):
            raise ValueError('Unrecognized option for highway_mode')
        if highway_mode != 'concat':
            raise NotImplementedError("only highway_mode='concat' is built (network.RenderingNet default, network.py:226)")
        self.submodule = submodule
        self.flag_outer = flag_outer
        self.gcn = gcn
        self.highway_mode = highway_mode
        common = dict(use_dropout=use_dropout, dropout_prob=dropout_prob, norm=norm)
        if gcn:
            # dead branch of the reference: parameters kept for checkpoint compatibility only
            self.fuse = DownBlock(inner_nc + out_channels_gcn, inner_nc, stride=1, kernal_size=3, **common)
        self.down = DownBlock(outer_nc, inner_nc, **common)
        self.up = UpBlock(2 * inner_nc if flag_outer else inner_nc, outer_nc, upsampling_mode=upsampling_mode, **common)

    forward = _holder_forward


class Unet(nn.Module):
    def __init__(self, in_channels, out_channels, nf0, num_down, max_channels, use_dropout, upsampling_mode='transpose',
                 dropout_prob=0.1, norm=nn.BatchNorm2d, outermost_linear=False, out_channels_gcn=512, use_gcn=True,
                 outermost_highway_mode='no_highway'):
        super().__init__()
        assert num_down > 0, "Need at least one downsampling layer in UNet."
        if norm is not nn.BatchNorm2d or not outermost_linear or not use_dropout or outermost_highway_mode != 'concat':
            raise NotImplementedError('the B200 engine builds the configuration network.RenderingNet uses: BatchNorm2d, '
                                      "dropout, linear out layer, outermost_highway_mode='concat' (network.py:236-247)")
        if num_down < 2:
            raise NotImplementedError('num_down >= 2 required')
        self.use_gcn = use_gcn
        self.outermost_highway_mode = outermost_highway_mode
        self.in_layer = nn.Sequential(Conv2dSame(in_channels, nf0, kernel_size=3, bias=False), norm(nf0, affine=True),
                                      nn.LeakyReLU(0.2, True), nn.Dropout2d(dropout_prob))
        ch = lambda i: min(2 ** i * nf0, max_channels)
        common = dict(use_dropout=use_dropout, dropout_prob=dropout_prob, upsampling_mode=upsampling_mode)
        block = UnetSkipConnectionBlock(ch(num_down - 1), ch(num_down - 1), norm=None, flag_outer=False, **common)
        for i in reversed(range(1, num_down - 1)):
            block = UnetSkipConnectionBlock(ch(i), ch(i + 1), submodule=block, norm=norm, **common)
        self.unet_block = UnetSkipConnectionBlock(ch(0), ch(1), submodule=block, norm=norm, gcn=use_gcn,
                                                  out_channels_gcn=out_channels_gcn, highway_mode=outermost_highway_mode, **common)
        self.out_layer = nn.Sequential(Conv2dSame(2 * nf0, out_channels, kernel_size=3, bias=True))
        self.out_layer_weight = self.out_layer[0].weight
        self._cfg = dict(in_channels=int(in_channels), out_channels=int(out_channels), nf0=int(nf0), num_down=int(num_down),
                         max_channels=int(max_channels), dropout_prob=float(dropout_prob))
        self._runner = UnetRunner(self)

    def forward(self, x, v_fea=None):
        """Pre-tanh output, as pytorch_prototyping.py:532-536 (``v_fea`` cannot influence it: SURVEY.md 3.4)."""
        return self._runner(x, apply_tanh=False)

    def forward_tanh(self, x):
        """tanh(Unet(x)) with the tanh fused into the last conv's epilogue (network.RenderingNet.forward)."""
        return self._runner(x, apply_tanh=True)

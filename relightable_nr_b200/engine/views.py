"""Strided 4-D (C, X, Y, N) views over channels-last tensors stored with a 1-pixel halo.

The generic convolution problem of include/rnr_b200.h addresses its A operand through such views;
this module builds them for the three access patterns of the U-Net
(pytorch_prototyping/pytorch_prototyping.py:112-115, 155-160, 242-264):

* ``padded``   : the whole [H+2, W+2] plane      -> 3x3/s1 convs after ReflectionPad2d(1)
* ``parity``   : every second row/col of the padded plane, phase (p, q) -> 4x4/s2 convs
* ``interior`` : the [H, W] plane without halo; out-of-range reads are zero -> ConvTranspose2d(p=1)
"""
import torch

from .._lib import View


class HaloTensor:
    """[N, H+2, W+2, C] 16-bit tensor.  Forward activations carry a reflect halo (written by the
    producing kernel); gradient tensors carry a zero halo (allocated zeroed, never written)."""

    def __init__(self, N, H, W, C, dtype, device, zero=True):
        self.N, self.H, self.W, self.C = N, H, W, C
        alloc = torch.zeros if zero else torch.empty
        self.t = alloc((N, H + 2, W + 2, C), dtype=dtype, device=device)

    @property
    def ptr(self):
        return self.t.data_ptr()

    @property
    def esize(self):
        return self.t.element_size()

    def _view(self, off_elems, dims, strides):
        v = View()
        v.ptr = self.ptr + off_elems * self.esize
        for i in range(4):
            v.dim[i] = dims[i]
            v.stride[i] = strides[i]
        return v

    def padded(self):
        C, Hp, Wp = self.C, self.H + 2, self.W + 2
        return self._view(0, (C, Wp, Hp, self.N), (1, C, Wp * C, Hp * Wp * C))

    def parity(self, p, q):
        """rows p, p+2, ... and cols q, q+2, ... of the padded plane."""
        C, Hp, Wp = self.C, self.H + 2, self.W + 2
        assert Hp % 2 == 0 and Wp % 2 == 0
        return self._view((p * Wp + q) * C, (C, Wp // 2, Hp // 2, self.N), (1, 2 * C, 2 * Wp * C, Hp * Wp * C))

    def interior(self):
        C, Hp, Wp = self.C, self.H + 2, self.W + 2
        return self._view((Wp + 1) * C, (C, self.W, self.H, self.N), (1, C, Wp * C, Hp * Wp * C))

    def interior_parity(self, p, q):
        """interior rows p, p+2, ... / cols q, q+2, ... (H, W even)."""
        C, Hp, Wp = self.C, self.H + 2, self.W + 2
        return self._view(((p + 1) * Wp + (q + 1)) * C, (C, self.W // 2, self.H // 2, self.N),
                          (1, 2 * C, 2 * Wp * C, Hp * Wp * C))

    def interior_tensor(self):
        return self.t[:, 1:-1, 1:-1, :]


def tile_shape(X):
    """(th, tw) with th*tw == 128 for an iteration space of width X."""
    tw = 16
    while tw > 1 and tw // 2 >= X:
        tw //= 2
    return 128 // tw, tw

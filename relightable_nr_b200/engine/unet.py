"""Host-side engine for the U-Net of network.RenderingNet (network.py:219-253,
pytorch_prototyping/pytorch_prototyping.py:432-536): forward, data-gradient and weight-gradient of
the 22 live convolution layers, each expressed as generic implicit-GEMM problems executed by
librnr_b200.so (tcgen05 kernels, or the SIMT validation kernels with impl='simt').

Data layout in HBM (per layer, allocated once per input shape and reused every step):
  act   fp16  [N, H+2, W+2, C]  reflect halo    -- post BN/activation/dropout, operand of the next conv
  raw   fp32  [N, H,   W,   C]                  -- conv output before BatchNorm (kept for backward)
  gz    bf16  [N, H+2, W+2, C]  zero halo       -- gradient w.r.t. raw
  gx    bf16  padded or dense                    -- gradient w.r.t. a layer's (padded) input
The dead GCN branch of UnetSkipConnectionBlock.forward (pytorch_prototyping.py:407-415, overwritten
at :416-419) is not executed; it cannot influence outputs or gradients.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from .. import _lib
from .._lib import BF16, EPI_BIAS, EPI_GSTATS, EPI_STATS, EPI_TANH, F16, F32, ConvProblem, GSrc, GStatSeg, KStep, WgradProblem, WPrepJob, WTap, WUnpackJob
from .views import HaloTensor, tile_shape

_TORCH_DT = {F16: torch.float16, BF16: torch.bfloat16, F32: torch.float32}


def _rup(x, m):
    return (x + m - 1) // m * m


@dataclass
class LayerSpec:
    name: str
    kind: str                  # 'c3' (3x3 s1 reflect) | 'c4s2' (4x4 s2 reflect) | 'ct' (ConvTranspose 4x4 s2 p1)
    src: List[str]             # names of input activation tensors (concatenated along channels)
    cin: List[int]             # channels of each source
    cout: int
    H: int                     # input spatial size
    W: int
    w_key: str
    b_key: Optional[str]
    bn_key: Optional[str]
    slope: Optional[float]     # LeakyReLU slope (0.0 == ReLU); None for the final linear layer (+tanh)
    dst: str
    drop: bool = True

    @property
    def Ho(self):
        return {'c3': self.H, 'c4s2': self.H // 2, 'ct': self.H * 2}[self.kind]

    @property
    def Wo(self):
        return {'c3': self.W, 'c4s2': self.W // 2, 'ct': self.W * 2}[self.kind]


def unet_layer_specs(in_channels, out_channels, nf0, num_down, max_channels, H, W) -> List[LayerSpec]:
    """Live layers of pytorch_prototyping.Unet in execution order.
    Keys are relative to the ``Unet`` module."""
    specs = []
    specs.append(LayerSpec('in', 'c3', ['input'], [in_channels], nf0, H, W,
                           'in_layer.0.net.1.weight', None, 'in_layer.1', 0.2, 'x0'))
    chans = [min(2 ** i * nf0, max_channels) for i in range(num_down)]
    up_specs = []
    prefix = 'unet_block'
    h, w = H, W
    for i in range(num_down):
        innermost = (i == num_down - 1)
        outer = chans[i]
        inner = chans[i] if innermost else chans[i + 1]
        bn = not innermost
        # DownBlock: pad, conv3, [bn], lrelu, drop, pad, conv4s2, [bn], lrelu, drop
        if bn:
            k1, b1, k2, b2 = 'net.1', 'net.2', 'net.6', 'net.7'
        else:
            k1, b1, k2, b2 = 'net.1', None, 'net.5', None
        specs.append(LayerSpec(f'b{i}.down1', 'c3', [f'x{i}'], [outer], outer, h, w,
                               f'{prefix}.down.{k1}.weight', None if bn else f'{prefix}.down.{k1}.bias',
                               f'{prefix}.down.{b1}' if bn else None, 0.2, f'd{i}'))
        specs.append(LayerSpec(f'b{i}.down2', 'c4s2', [f'd{i}'], [outer], inner, h, w,
                               f'{prefix}.down.{k2}.weight', None if bn else f'{prefix}.down.{k2}.bias',
                               f'{prefix}.down.{b2}' if bn else None, 0.2, f'x{i + 1}'))
        # UpBlock: convT, [bn], relu, drop, Conv2dSame, [bn], relu, drop
        if bn:
            u1, ub1, u2, ub2 = 'net.0', 'net.1', 'net.4.net.1', 'net.5'
        else:
            u1, ub1, u2, ub2 = 'net.0', None, 'net.3.net.1', None
        if innermost:
            srcs, cins = [f'x{i + 1}'], [inner]
        else:
            srcs, cins = [f'x{i + 1}', f'y{i + 1}'], [inner, inner]
        up_specs.append([
            LayerSpec(f'b{i}.up1', 'ct', srcs, cins, outer, h // 2, w // 2,
                      f'{prefix}.up.{u1}.weight', None if bn else f'{prefix}.up.{u1}.bias',
                      f'{prefix}.up.{ub1}' if bn else None, 0.0, f'u{i}'),
            LayerSpec(f'b{i}.up2', 'c3', [f'u{i}'], [outer], outer, h, w,
                      f'{prefix}.up.{u2}.weight', None if bn else f'{prefix}.up.{u2}.bias',
                      f'{prefix}.up.{ub2}' if bn else None, 0.0, f'y{i}'),
        ])
        prefix += '.submodule'
        h //= 2
        w //= 2
    for ups in reversed(up_specs):
        specs.extend(ups)
    specs.append(LayerSpec('out', 'c3', ['x0', 'y0'], [nf0, nf0], out_channels, H, W,
                           'out_layer.0.net.1.weight', 'out_layer.0.net.1.bias', None, None, 'out', drop=False))
    return specs


class _Plan:
    """RAII holder of a C-side plan handle."""

    def __init__(self, handle, kind):
        self.h = handle
        self.kind = kind

    def __del__(self):
        try:
            L = _lib.lib()
            if self.h:
                {'conv': L.rnr_conv_plan_destroy, 'wgrad': L.rnr_wgrad_plan_destroy, 'wprep': L.rnr_wprep_plan_destroy,
                 'wunpack': L.rnr_wgrad_unpack_plan_destroy}[self.kind](self.h)
        except Exception:
            pass


@dataclass
class _WPrep:
    src_key: str
    src_off: int            # element offset into the fp32 parameter
    dst: torch.Tensor
    dtype: int
    nr: int
    nr_pad: int
    nc: int
    cpad: int
    ntaps: int
    s_r: int
    s_c: int
    tapoff: torch.Tensor    # int32 device
    tapoff_host: Tuple[int, ...] = ()
    chunked: int = 0        # 1: Wmat columns are [chunk of 64 channels][tap][64] (halo-kernel order)
    ld: int = 0             # destination row pitch in elements (0: ntaps * cpad) -- a job may fill one source's column block


@dataclass
class _LayerState:
    spec: LayerSpec
    raw: Optional[torch.Tensor] = None
    stats: Optional[torch.Tensor] = None
    n_stat_tiles: int = 0
    mean: Optional[torch.Tensor] = None
    invstd: Optional[torch.Tensor] = None
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    c1: Optional[torch.Tensor] = None
    c2: Optional[torch.Tensor] = None
    coef: Optional[torch.Tensor] = None
    bwd_partials: Optional[torch.Tensor] = None
    fwd_plans: List[_Plan] = field(default_factory=list)
    dgrad_plans: List[_Plan] = field(default_factory=list)
    wgrad_plan: Optional[_Plan] = None
    wprep_fwd: List[_WPrep] = field(default_factory=list)
    wprep_dgrad: List[_WPrep] = field(default_factory=list)
    gx: Optional[torch.Tensor] = None      # dgrad result
    gx_fold: bool = False
    gx_ld: int = 0
    drop: Optional[torch.Tensor] = None
    ticket: Optional[torch.Tensor] = None
    bwd_totals: Optional[torch.Tensor] = None
    wmat_fwd: Optional[dict] = None            # layout of the forward GEMM matrix family (for the optimiser pass that re-derives it)
    wmat_dgrad: Optional[dict] = None          # same for the data-gradient family
    bn_ticket: Optional[torch.Tensor] = None   # last-CTA ticket of the fused forward BatchNorm finalize
    fwd_totals: Optional[torch.Tensor] = None  # [2, C] fp64 batch sums written by the conv launch (no finalize launch)
    fwd_ticket: Optional[torch.Tensor] = None
    bn_fused: bool = False                     # statistics finalize runs inside the conv kernel (rnr_conv_plan_set_bn)


class UNetEngine:
    def __init__(self, specs: List[LayerSpec], params: Dict[str, torch.Tensor], buffers: Dict[str, torch.Tensor],
                 N: int, in_channels: int, device, impl: str = 'tc', input_grad_range: Optional[Tuple[int, int]] = None,
                 act_dtype: int = F16, grad_dtype: int = BF16, need_backward: bool = True, wgrad_impl: Optional[str] = None,
                 final_tanh: bool = True):
        self.L = _lib.lib()
        self.specs = specs
        self.params = params
        self.buffers = buffers
        self.N = N
        self.device = torch.device(device)
        self.impl = {'simt': 0, 'tc': 1}[impl]
        self.wgrad_impl = {'simt': 0, 'tc': 1}[wgrad_impl or impl]
        self.act_dt, self.grad_dt = act_dtype, grad_dtype
        # pre-BatchNorm convolution outputs: stored in the activation format on the tensor-core path (the statistics come from
        # the fp32 accumulators in the epilogue; every BN pass then streams 2 instead of 4 bytes per element), fp32 on the SIMT
        # validation path.  RNR_RAW_FP32=1 keeps fp32.
        self.raw_dt = act_dtype if (impl == 'tc' and os.environ.get('RNR_RAW_FP32', '0') != '1') else F32
        self.in_channels = in_channels
        self.in_cpad = _rup(in_channels, 64) if in_channels > 32 else _rup(in_channels, 16)
        self.input_grad_range = input_grad_range
        self.need_backward = need_backward
        self.final_tanh = final_tanh
        self._zeros_out = None
        self.H, self.W = specs[0].H, specs[0].W
        self.keep = []            # keeps ctypes arrays / tensors referenced by plans alive
        self.gpu_launches = 0
        self.timing = None        # list of (kind, layer, start_event, end_event) when enabled (bench.py roofline leg)
        self._bn_mode = None      # (training, momentum, eps) the fused finalize is currently configured for
        self._gi = None           # persistent input-gradient buffer of the fused step
        self._build()

    # ------------------------------------------------------------------------------------------
    # construction
    # ------------------------------------------------------------------------------------------
    def _alloc(self, shape, dtype, zero=False):
        f = torch.zeros if zero else torch.empty
        return f(shape, dtype=dtype, device=self.device)

    def _build(self):
        N, dev = self.N, self.device
        adt, gdt = _TORCH_DT[self.act_dt], _TORCH_DT[self.grad_dt]
        self.acts: Dict[str, HaloTensor] = {}
        self.gz: Dict[str, HaloTensor] = {}
        self.layers: Dict[str, _LayerState] = {}
        self.producer: Dict[str, str] = {}        # act name -> layer name producing it
        self.consumers: Dict[str, List[Tuple[str, int]]] = {}   # act name -> [(layer name, source index)]
        s0 = self.specs[0]
        self.acts['input'] = HaloTensor(N, s0.H, s0.W, self.in_cpad, adt, dev, zero=True)
        # bf16 copies of the activations: B operand of the weight-gradient MMA (same format as the gradients)
        self.acts_w: Dict[str, HaloTensor] = {}
        # The weight-gradient MMA multiplies bf16 gradients with the activations.  kind::f16 with A = fp16 and B = bf16 (or the
        # reverse) raises "illegal instruction" on sm_100a (measured: rnr_wgrad_run with RNR_DUAL=0), so the forward pass keeps
        # a bf16 copy of every activation next to the fp16 one.
        self.dual = self.need_backward and self.act_dt != self.grad_dt and os.environ.get('RNR_DUAL', '1') != '0'
        if self.dual:
            self.acts_w['input'] = HaloTensor(N, s0.H, s0.W, self.in_cpad, gdt, dev, zero=True)
        self.ones = {}
        self.zeros = {}
        grad_numel = 0
        self.grad_slices = {}
        for sp in self.specs:
            st = _LayerState(sp)
            self.layers[sp.name] = st
            for si, s in enumerate(sp.src):
                self.consumers.setdefault(s, []).append((sp.name, si))
            self.producer[sp.dst] = sp.name
            Ho, Wo = sp.Ho, sp.Wo
            if sp.dst == 'out':
                # channel pitch of the final layer's output / gradient: a multiple of 64 keeps its data-gradient on the halo kernel
                self.out_ld = _rup(sp.cout, 64) if sp.cout > 32 else _rup(sp.cout, 16)
                st.raw = self._alloc((N, Ho, Wo, self.out_ld), torch.float32, zero=True)
            else:
                st.raw = self._alloc((N, Ho, Wo, sp.cout), _TORCH_DT[self.raw_dt])
                self.acts[sp.dst] = HaloTensor(N, Ho, Wo, sp.cout, adt, dev, zero=True)
                if self.dual:
                    self.acts_w[sp.dst] = HaloTensor(N, Ho, Wo, sp.cout, gdt, dev, zero=True)
                for nm in ('mean', 'invstd', 'scale', 'shift', 'c1', 'c2'):
                    setattr(st, nm, self._alloc((sp.cout,), torch.float32, zero=True))
                st.coef = self._alloc((3 * sp.cout,), torch.float32, zero=True)
                st.invstd.fill_(1.0)
                st.scale.fill_(1.0)
            if self.need_backward:
                ld = self.out_ld if sp.dst == 'out' else sp.cout
                self.gz[sp.name] = HaloTensor(N, Ho, Wo, ld, gdt, dev, zero=True)
                st.bwd_partials = self._alloc((148 * 8 * 2 * max(ld, 8),), torch.float32)
                st.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
                st.bwd_totals = torch.zeros(2 * max(ld, 8), dtype=torch.float64, device=dev)
        # flat gradient storage: all convolution weights first, then the small tensors (biases, BatchNorm affine) as one
        # contiguous tail -- a data-parallel step all-reduces the weight gradients early, in GEMM order (wscratch), and only this
        # tail after the backward pass
        for sp in self.specs:
            n = self.params[sp.w_key].numel()
            self.grad_slices[sp.w_key] = (grad_numel, n)
            grad_numel += _rup(n, 4)
        self.grad_small_offset = grad_numel
        for sp in self.specs:
            for key in (sp.b_key, (sp.bn_key + '.weight') if sp.bn_key else None, (sp.bn_key + '.bias') if sp.bn_key else None):
                if key is not None:
                    n = self.params[key].numel()
                    self.grad_slices[key] = (grad_numel, n)
                    grad_numel += _rup(n, 4)
        self.grad_flat = self._alloc((max(grad_numel, 4),), torch.float32, zero=True) if self.need_backward else None
        # weight gradients are accumulated in GEMM order [tap][co][ci] (vector reductions, see csrc/wprep.cu: wgrad_unpack) and
        # moved into grad_flat by one batched launch at the end of the backward pass
        self.wscratch = None
        self.wscratch_slices = {}
        self._wunpack_jobs = []
        self._wunpack_names = []
        if self.need_backward and self.wgrad_impl == 1:
            n = 0
            for sp in self.specs:
                if sum(sp.cin) % 4 == 0:
                    k = 9 if sp.kind == 'c3' else 16
                    self.wscratch_slices[sp.name] = (n, k * sp.cout * sum(sp.cin))
                    n += _rup(k * sp.cout * sum(sp.cin), 4)
            if n:
                self.wscratch = self._alloc((n,), torch.float32, zero=True)
        for sp in self.specs:
            self._build_layer(self.layers[sp.name])
        self.wunpack_plan = None
        if self._wunpack_jobs:
            arr = (WUnpackJob * len(self._wunpack_jobs))()
            for j, (src, dst, cout, cin, ntaps, s_co, s_ci) in zip(arr, self._wunpack_jobs):
                j.src, j.dst, j.cout, j.cin, j.ntaps, j.s_co, j.s_ci = src, dst, cout, cin, ntaps, s_co, s_ci
            h = C.c_void_p()
            _lib.check(self.L.rnr_wgrad_unpack_plan_create(arr, len(self._wunpack_jobs), C.byref(h)), 'rnr_wgrad_unpack_plan_create')
            self.wunpack_plan = _Plan(h, 'wunpack')
        # BatchNorm layers whose backward statistics come out of their consumers' data-gradient launches (RNR_BN_BWD_FUSED=1).  OFF by
        # default: measured on B200 with the CTA-pair conv kernel, alternating on one box, 216.4 / 216.8 views/s with it and 216.3 /
        # 215.8 without -- the 0.25 ms of reduction passes it removes come back as +0.16 ms in the data-gradient epilogues
        # (profiles/r02_perf_unet_c23_fused{0,1}.txt), and the dominant kernel's roofline fraction drops from 0.30 to 0.27.
        self.gstat_layers = set()
        self._gstat_keys = {}
        if self.need_backward and self.impl == 1 and os.environ.get('RNR_BN_BWD_FUSED', '0') == '1':
            self._plan_gstats()
        fwd_items = [w for sp in self.specs for w in self.layers[sp.name].wprep_fwd]
        all_items = fwd_items + [w for sp in self.specs for w in self.layers[sp.name].wprep_dgrad]
        self.wprep_fwd_plan = self._wprep_plan(fwd_items)
        self.wprep_all_plan = self._wprep_plan(all_items) if len(all_items) > len(fwd_items) else self.wprep_fwd_plan
        self.wprep_dgrad_plan = self._wprep_plan(all_items[len(fwd_items):]) if len(all_items) > len(fwd_items) else None

    # ---- BatchNorm-backward statistics inside the data-gradient launches ------------------------
    def _gstat_segments(self, consumer, fused):
        """rnr_gstat_seg_t array of one consumer layer's data-gradient plan: one segment per concatenated input; a segment is live
        when its producer layer is in ``fused``."""
        sp = self.layers[consumer].spec
        r0 = self.input_grad_range[0] if (consumer == 'in' and self.input_grad_range is not None) else 0
        segs = (GStatSeg * len(sp.src))()
        c0 = 0
        for si, src in enumerate(sp.src):
            d = segs[si]
            d.c_lo, d.c_hi = c0 - r0, c0 + sp.cin[si] - r0
            c0 += sp.cin[si]
            prod = self._producer.get(src)
            d.raw = None
            if prod is not None and prod in fused:
                pst = self.layers[prod]
                d.raw, d.raw_dtype, d.C = pst.raw.data_ptr(), self.raw_dt, pst.spec.cout
                d.scale, d.shift, d.mean = pst.scale.data_ptr(), pst.shift.data_ptr(), pst.mean.data_ptr()
                d.drop = pst.drop.data_ptr() if pst.drop is not None else None
                d.slope, d.totals = pst.spec.slope, pst.bwd_totals.data_ptr()
        return segs

    def _set_gstats(self, consumer, fused):
        st = self.layers[consumer]
        sp = st.spec
        segs = self._gstat_segments(consumer, fused)
        if not any(s.raw for s in segs):
            return self.L.rnr_conv_plan_set_gstats(st.dgrad_plans[0].h, segs, 0, sp.H, sp.W, 0) if st.dgrad_plans else 0
        return self.L.rnr_conv_plan_set_gstats(st.dgrad_plans[0].h, segs, len(sp.src), sp.H, sp.W, 1 if st.gx_fold else 0)

    def _plan_gstats(self):
        """Decide which BatchNorm layers get their backward sums from the data-gradient epilogues of their consumers (every consumer
        must be ONE halo-kernel launch that accepts the segment layout), then configure those plans.  The rest keep the separate
        reduction pass (rnr_bn_bwd_reduce_fin)."""
        self._producer = {sp.dst: sp.name for sp in self.specs}
        fused = set()
        for sp in self.specs:
            vpp = sp.cout // 8
            if (sp.bn_key is None or sp.dst == 'out' or self.raw_dt == F32 or sp.cout % 8 or vpp & (vpp - 1) or vpp > 256):
                continue
            cons = self.consumers.get(sp.dst, [])
            if cons and all(self.layers[ln].gx is not None and len(self.layers[ln].dgrad_plans) == 1 for ln, _ in cons):
                fused.add(sp.name)
        changed = True
        while changed and fused:
            changed = False
            for sp in self.specs:
                st = self.layers[sp.name]
                if st.gx is None or len(st.dgrad_plans) != 1:
                    continue
                prods = {self._producer.get(s) for s in sp.src} & fused
                if prods and self._set_gstats(sp.name, fused) != 0:
                    fused -= prods          # this launch cannot carry the statistics: its producers fall back (and so must every
                    changed = True          # other consumer of theirs -- reconfigure from the top)
                    break
        for sp in self.specs:
            st = self.layers[sp.name]
            if st.gx is not None and len(st.dgrad_plans) == 1:
                rc = self._set_gstats(sp.name, fused)
                assert rc == 0 or not ({self._producer.get(s) for s in sp.src} & fused)
        self.gstat_layers = fused
        self._gstat_keys = {}

    def _refresh_gstats(self):
        """The segments hold the producers' dropout-mask pointers: re-arm a plan when a forward pass installed other masks."""
        if not self.gstat_layers:
            return
        for sp in self.specs:
            prods = [self._producer.get(s) for s in sp.src]
            if not any(pn in self.gstat_layers for pn in prods):
                continue
            key = tuple(self.layers[pn].drop.data_ptr() if (pn in self.gstat_layers and self.layers[pn].drop is not None) else 0
                        for pn in prods)
            if self._gstat_keys.get(sp.name, tuple(0 for _ in prods)) != key:
                _lib.check(self._set_gstats(sp.name, self.gstat_layers), 'rnr_conv_plan_set_gstats(%s)' % sp.name)
                self._gstat_keys[sp.name] = key

    # ---- problem builders ---------------------------------------------------------------------
    def _conv_problem(self, views, ksteps, ab_dtype, bk, wmat, n_rows_w, cout, mN, mY, mX, out_t, out_dtype,
                      out_strides, out_mp, epi, bias, stats, ldstats, impl, defer=False):
        prob = ConvProblem()
        for i, v in enumerate(views):
            prob.views[i] = v
        prob.n_views = len(views)
        prob.ab_dtype = ab_dtype
        prob.bk = bk
        arr = (KStep * len(ksteps))()
        for i, (v, c0, dx, dy) in enumerate(ksteps):
            arr[i].view, arr[i].c0, arr[i].dx, arr[i].dy = v, c0, dx, dy
        prob.n_ksteps = len(ksteps)
        prob.ksteps = C.cast(arr, C.POINTER(KStep))
        prob.wmat = wmat if isinstance(wmat, int) else wmat.data_ptr()
        prob.n_rows_w = n_rows_w
        prob.cout = cout
        prob.mN, prob.mY, prob.mX = mN, mY, mX
        prob.th, prob.tw = tile_shape(mX)
        prob.out = out_t if isinstance(out_t, int) else out_t.data_ptr()
        prob.out_dtype = out_dtype
        prob.out_sn, prob.out_sy, prob.out_sx = out_strides
        prob.out_my, prob.out_mx, prob.out_py, prob.out_px = out_mp
        prob.epi = epi
        prob.bias = bias.data_ptr() if bias is not None else None
        prob.stats = stats if isinstance(stats, int) or stats is None else stats.data_ptr()
        prob.ldstats = ldstats
        prob._keep = arr
        if defer:
            return prob
        h = C.c_void_p()
        _lib.check(self.L.rnr_conv_plan_create(C.byref(prob), impl, C.byref(h)), 'rnr_conv_plan_create')
        return _Plan(h, 'conv')

    def _fused_plan(self, probs, impl):
        """One plan (one launch) for sub-problems that differ only in taps / weights / output parity; None when the library
        cannot fuse them (SIMT validation path, per-tap kernel)."""
        if impl != 1 or len(probs) < 2:
            return None
        arr = (ConvProblem * len(probs))()
        for i, pr in enumerate(probs):
            C.memmove(C.byref(arr[i]), C.byref(pr), C.sizeof(ConvProblem))
        h = C.c_void_p()
        rc = self.L.rnr_conv_plan_create_multi(arr, len(probs), impl, C.byref(h))
        if rc != 0:
            return None
        return _Plan(h, 'conv')

    def _single_plan(self, prob, impl):
        h = C.c_void_p()
        _lib.check(self.L.rnr_conv_plan_create(C.byref(prob), impl, C.byref(h)), 'rnr_conv_plan_create')
        return _Plan(h, 'conv')

    @staticmethod
    def _bk_for(channel_counts):
        for bk in (64, 32, 16):
            if all(c % bk == 0 for c in channel_counts):
                return bk
        raise ValueError('channel counts %s must be multiples of 16' % (channel_counts,))

    @staticmethod
    def _order_ksteps(taps, C, bk, chunked):
        """K-steps of a single-source problem: taps = [(view, dx, dy)]; chunk-major when ``chunked`` else tap-major."""
        if chunked:
            return [(v, c0, dx, dy) for c0 in range(0, C, bk) for (v, dx, dy) in taps]
        return [(v, c0, dx, dy) for (v, dx, dy) in taps for c0 in range(0, C, bk)]

    def _tapoff(self, offs):
        t = torch.tensor(offs, dtype=torch.int32, device=self.device)
        t.host = tuple(int(o) for o in offs)
        self.keep.append(t)
        return t

    def _wprep_plan(self, items):
        """One batched weight-preparation plan (a single launch) for a list of _WPrep jobs."""
        arr = (WPrepJob * len(items))()
        for j, w in zip(arr, items):
            src = self.params[w.src_key]
            j.src = src.data_ptr() + 4 * w.src_off
            j.dst = w.dst.data_ptr()
            j.dst_dtype, j.nr, j.nr_pad, j.nc, j.cpad, j.ntaps = w.dtype, w.nr, w.nr_pad, w.nc, w.cpad, w.ntaps
            j.s_r, j.s_c = w.s_r, w.s_c
            j.chunked = w.chunked
            j.ld = w.ld
            for t, o in enumerate(w.tapoff.host):
                j.tapoff[t] = o
        h = C.c_void_p()
        _lib.check(self.L.rnr_wprep_plan_create(arr, len(items), C.byref(h)), 'rnr_wprep_plan_create')
        return _Plan(h, 'wprep')

    def _build_layer(self, st: _LayerState):
        sp = st.spec
        N = self.N
        adt_t, gdt_t = _TORCH_DT[self.act_dt], _TORCH_DT[self.grad_dt]
        srcs = [self.acts[s] for s in sp.src]
        cpads = [t.C for t in srcs]           # stored channel counts (input layer is padded)
        # Channel counts that are not multiples of 64 (DNR at nf0 = 80: 80 / 160; the 16-channel DNR input; test-sized nets): the
        # K loop still runs in 64-channel chunks -- the TMA box simply reaches past the tensor's last channel and the hardware
        # zero-fills what is out of bounds, so the activations stay UNPADDED in HBM and every layer runs on the halo kernel.  Only
        # the weight matrix carries the (zero) padding columns.  RNR_CONV_OOB=0 restores exact 16 / 32-channel chunks (per-tap kernel).
        oob = (self.impl == 1 and any(c % 64 for c in cpads) and all(c % 8 == 0 for c in cpads)
               and os.environ.get('RNR_CONV_OOB', '1') != '0')
        kpads = [_rup(c, 64) for c in cpads] if oob else list(cpads)
        cin_tot = sum(sp.cin)
        cpad_tot = sum(kpads)
        final = sp.dst == 'out'
        Ho, Wo = sp.Ho, sp.Wo
        cout = sp.cout
        ld_out = self.out_ld if final else cout
        k = 3 if sp.kind == 'c3' else 4
        kk = k * k
        bk = 64 if oob else self._bk_for(cpads)
        n_rows = _rup(cout, 16)
        epi = 0
        bias_t = None
        if final:
            epi = EPI_BIAS | (EPI_TANH if self.final_tanh else 0)
            bias_t = self.params[sp.b_key]
        elif sp.bn_key is not None:
            epi = EPI_STATS

        chunked = 1 if bk == 64 else 0

        def ksteps_for(taps):
            """taps: list of (view index per source -> list, dx, dy).  K-step j multiplies Wmat columns [j*bk, (j+1)*bk).
            bk == 64: chunk-major order (all taps of a 64-channel chunk adjacent -- the halo kernel loads that chunk's
            input tile once and walks the taps inside shared memory); otherwise tap-major."""
            ks = []
            if chunked:
                for si in range(len(srcs)):
                    for c0 in range(0, kpads[si], bk):
                        for (vidx, dx, dy) in taps:
                            ks.append((vidx[si], c0, dx, dy))
                return ks
            for (vidx, dx, dy) in taps:
                for si in range(len(srcs)):
                    for c0 in range(0, kpads[si], bk):
                        ks.append((vidx[si], c0, dx, dy))
            return ks

        def fwd_wprep(wm, ntaps, s_r, s_c, tapoffs):
            """Weight-preparation job(s) of one forward matrix ``wm`` [rows, ntaps * cpad_tot].  With out-of-bounds chunks every
            source owns its own (padded) column block [chunks of the source][tap][64]: one job per source, row pitch = the matrix."""
            toff = self._tapoff(tapoffs)
            if not oob or len(srcs) == 1:
                return [_WPrep(sp.w_key, 0, wm, self.act_dt, cout, n_rows, cin_tot, cpad_tot, ntaps, s_r, s_c, toff, chunked=chunked)]
            jobs, ci0, col0 = [], 0, 0
            for si in range(len(srcs)):
                jobs.append(_WPrep(sp.w_key, ci0 * s_c, wm[:, col0:], self.act_dt, cout, n_rows, sp.cin[si], kpads[si], ntaps, s_r, s_c, toff,
                                   chunked=chunked, ld=ntaps * cpad_tot))
                ci0 += sp.cin[si]
                col0 += ntaps * kpads[si]
            return jobs

        def wmat_desc(base, ld, dtype, ntaps, sub_rows, sub_taps, r0=0, r1=0, ok=None):
            """Layout record of a GEMM matrix family for the fused optimiser pass (csrc/optim.cu): source tap t sits at slot k of
            sub-matrix s with inv[t] = s*16 + k.  None when the layout is outside what that pass writes (tap-major columns, or
            several concatenated sources with out-of-bounds padding between them)."""
            supported = chunked and (not oob or len(srcs) == 1) if ok is None else ok
            if not supported:
                return None
            inv = [-1] * 16
            for s_i, offs in enumerate(sub_taps):
                for k_i, t in enumerate(offs):
                    inv[t] = s_i * 16 + k_i
            return dict(base=base, ld=int(ld), dtype=dtype, ntaps=int(ntaps), sub_rows=int(sub_rows), r0=int(r0), r1=int(r1), inv=inv)

        # ------------------------------ forward ------------------------------
        if sp.kind == 'c3':
            views = [t.padded() for t in srcs]
            taps = [([si for si in range(len(srcs))], kw, kh) for kh in range(3) for kw in range(3)]
            tapoffs = [kh * 3 + kw for kh in range(3) for kw in range(3)]
            wm = self._alloc((n_rows, 9 * cpad_tot), adt_t, zero=True)
            st.wprep_fwd += fwd_wprep(wm, 9, cin_tot * 9, 9, tapoffs)
            st.wmat_fwd = wmat_desc(wm, 9 * cpad_tot, self.act_dt, 9, n_rows, [tapoffs])
            th, tw = tile_shape(Wo)
            tiles = max(N * -(-Ho // th) * -(-Wo // tw), 148)      # rows of the partial-sum buffer: one per tile or per CTA
            if epi & EPI_STATS:
                st.stats = self._alloc((tiles, 2, cout), torch.float32, zero=True)
            st.fwd_plans.append(self._conv_problem(
                views, ksteps_for(taps), self.act_dt, bk, wm, n_rows, cout, N, Ho, Wo, st.raw, F32 if final else self.raw_dt,
                (Ho * Wo * ld_out, Wo * ld_out, ld_out), (1, 1, 0, 0), epi, bias_t, st.stats, cout, self.impl))
            st.n_stat_tiles = self.L.rnr_conv_plan_stat_rows(st.fwd_plans[-1].h)
        elif sp.kind == 'c4s2':
            views = []
            for t in srcs:
                views += [t.parity(p, q) for p in range(2) for q in range(2)]
            taps, tapoffs = [], []
            for p in range(2):                      # parity-major: the 4 taps reading one parity view are adjacent
                for q in range(2):
                    for a in range(2):
                        for b in range(2):
                            kh, kw = 2 * a + p, 2 * b + q
                            taps.append(([si * 4 + p * 2 + q for si in range(len(srcs))], b, a))
                            tapoffs.append(kh * 4 + kw)
            wm = self._alloc((n_rows, 16 * cpad_tot), adt_t, zero=True)
            st.wprep_fwd += fwd_wprep(wm, 16, cin_tot * 16, 16, tapoffs)
            st.wmat_fwd = wmat_desc(wm, 16 * cpad_tot, self.act_dt, 16, n_rows, [tapoffs])
            th, tw = tile_shape(Wo)
            tiles = max(N * -(-Ho // th) * -(-Wo // tw), 148)
            if epi & EPI_STATS:
                st.stats = self._alloc((tiles, 2, cout), torch.float32, zero=True)
            st.fwd_plans.append(self._conv_problem(
                views, ksteps_for(taps), self.act_dt, bk, wm, n_rows, cout, N, Ho, Wo, st.raw, F32 if final else self.raw_dt,
                (Ho * Wo * ld_out, Wo * ld_out, ld_out), (1, 1, 0, 0), epi, bias_t, st.stats, cout, self.impl))
            st.n_stat_tiles = self.L.rnr_conv_plan_stat_rows(st.fwd_plans[-1].h)
        else:  # ConvTranspose 4x4 s2 p1: four parity sub-problems, each a 2x2 conv over the un-padded input
            views = [t.interior() for t in srcs]
            Hi, Wi = sp.H, sp.W
            th, tw = tile_shape(Wi)
            tiles = max(N * -(-Hi // th) * -(-Wi // tw), 148)      # rows reserved per parity sub-problem (unused rows stay zero)
            if epi & EPI_STATS:
                st.stats = self._alloc((4 * tiles, 2, cout), torch.float32, zero=True)
                st.n_stat_tiles = 4 * tiles
            sel = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}      # parity -> [(k index, input offset)]
            # the four parity sub-problems share one weight buffer (sub s at rows [s*n_rows, (s+1)*n_rows)) so that they can
            # run as ONE launch (rnr_conv_plan_create_multi): 4x the tiles per launch instead of four half-empty grids
            wm_all = self._alloc((4 * n_rows, 4 * cpad_tot), adt_t, zero=True)
            probs = []
            sub_taps = []
            for ph in range(2):
                for pw in range(2):
                    taps, tapoffs = [], []
                    for (kh, dy) in sel[ph]:
                        for (kw, dx) in sel[pw]:
                            taps.append(([si for si in range(len(srcs))], dx, dy))
                            tapoffs.append(kh * 4 + kw)
                    sub_taps.append(tapoffs)
                    sidx = ph * 2 + pw
                    wm = wm_all[sidx * n_rows:(sidx + 1) * n_rows]
                    # weight [Cin, Cout, 4, 4]: rows = co (stride 16), cols = ci (stride Cout*16)
                    st.wprep_fwd += fwd_wprep(wm, 4, 16, cout * 16, tapoffs)
                    probs.append(self._conv_problem(
                        views, ksteps_for(taps), self.act_dt, bk, wm, n_rows, cout, N, Hi, Wi, st.raw, F32 if final else self.raw_dt,
                        (Ho * Wo * ld_out, Wo * ld_out, ld_out), (2, 2, ph, pw), epi, bias_t, st.stats, cout, self.impl, defer=True))
            st.wmat_fwd = wmat_desc(wm_all, 4 * cpad_tot, self.act_dt, 4, n_rows, sub_taps)
            self.keep.append(probs)
            fused = self._fused_plan(probs, self.impl)
            if fused is not None:
                st.fwd_plans.append(fused)
                if st.stats is not None:
                    st.n_stat_tiles = self.L.rnr_conv_plan_stat_rows(fused.h)
            else:
                rows_pp = 0
                for sidx, pr in enumerate(probs):
                    if st.stats is not None:
                        pr.stats = st.stats.data_ptr() + sidx * rows_pp * 2 * cout * 4
                    st.fwd_plans.append(self._single_plan(pr, self.impl))
                    if sidx == 0:
                        # the four parity sub-problems have identical shapes: each writes `rows_pp` consecutive partial-sum rows
                        rows_pp = self.L.rnr_conv_plan_stat_rows(st.fwd_plans[-1].h)
                        if st.stats is not None:
                            st.n_stat_tiles = 4 * rows_pp
        # BatchNorm finalize inside the conv kernel's last CTA: possible when the layer is ONE halo-kernel launch.  OFF by default:
        # measured on B200 (bench.py, 512^2 step) 198.5 views/s with it vs 205.1 without -- one CTA reducing 148 rows while 147 SMs
        # idle at the tail of every conv launch costs more than the 17 tiny finalize launches it removes (RNR_BN_FUSED_FINALIZE=1 enables)
        if (epi & EPI_STATS) and len(st.fwd_plans) == 1 and self.impl == 1 and os.environ.get('RNR_BN_FUSED_FINALIZE', '0') == '1':
            st.bn_ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
            st.bn_fused = self._set_bn(st, True, 0.1, 1e-5, probe=True)
        # BatchNorm batch sums as fp64 totals consumed by the activation pass itself (rnr_bn_act_fwd_tot): no finalize launch.  OFF by
        # default: measured on B200, alternating on one box, 218.3 / 218.4 views/s with it vs 222.8 / 221.3 without -- the per-block
        # coefficient prologue (an L2 round trip for the totals, fp64 math, one ticket atomic) of ~2000 blocks per layer costs more
        # than the 17 tiny finalize launches it removes (RNR_BN_FWD_TOTALS=1 enables; parity-tested)
        if ((epi & EPI_STATS) and len(st.fwd_plans) == 1 and self.impl == 1 and not st.bn_fused and
                os.environ.get('RNR_BN_FWD_TOTALS', '0') == '1'):
            tot = torch.zeros(2 * cout, dtype=torch.float64, device=self.device)
            if self.L.rnr_conv_plan_set_stat_totals(st.fwd_plans[0].h, tot.data_ptr()) == 0:
                st.fwd_totals = tot
                st.fwd_ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        if not self.need_backward:
            return

        # ------------------------------ weight gradient ------------------------------
        G = self.gz[sp.name]
        wp = WgradProblem()
        if self.dual:
            wsrcs = [self.acts_w[s] for s in sp.src]
            if sp.kind == 'c3':
                wviews = [t.padded() for t in wsrcs]
            elif sp.kind == 'c4s2':
                wviews = []
                for t in wsrcs:
                    wviews += [t.parity(p, q) for p in range(2) for q in range(2)]
            else:
                wviews = [t.interior() for t in wsrcs]
            w_dt = self.grad_dt
        else:
            wviews, w_dt = views, self.act_dt
        for i, v in enumerate(wviews):
            wp.aviews[i] = v
        wp.n_aviews = len(wviews)
        wp.a_dtype, wp.g_dtype = w_dt, self.grad_dt
        wtaps = []
        if sp.kind == 'c3':
            wp.gviews[0] = G.interior()
            wp.n_gviews = 1
            for kh in range(3):
                for kw in range(3):
                    ci0 = 0
                    for si in range(len(srcs)):
                        wtaps.append((si, kw, kh, 0, 0, ci0, sp.cin[si], kh * 3 + kw))
                        ci0 += sp.cin[si]
            wp.mN, wp.mY, wp.mX = N, Ho, Wo
            wp.s_co, wp.s_ci = cin_tot * 9, 9
        elif sp.kind == 'c4s2':
            wp.gviews[0] = G.interior()
            wp.n_gviews = 1
            for kh in range(4):
                for kw in range(4):
                    a, p, b, q = kh // 2, kh % 2, kw // 2, kw % 2
                    ci0 = 0
                    for si in range(len(srcs)):
                        wtaps.append((si * 4 + p * 2 + q, b, a, 0, 0, ci0, sp.cin[si], kh * 4 + kw))
                        ci0 += sp.cin[si]
            wp.mN, wp.mY, wp.mX = N, Ho, Wo
            wp.s_co, wp.s_ci = cin_tot * 16, 16
        else:
            sel = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}
            for ph in range(2):
                for pw in range(2):
                    wp.gviews[ph * 2 + pw] = G.interior_parity(ph, pw)
                    for (kh, dy) in sel[ph]:
                        for (kw, dx) in sel[pw]:
                            ci0 = 0
                            for si in range(len(srcs)):
                                wtaps.append((si, dx, dy, ph * 2 + pw, 0, ci0, sp.cin[si], kh * 4 + kw))
                                ci0 += sp.cin[si]
            wp.n_gviews = 4
            wp.mN, wp.mY, wp.mX = N, sp.H, sp.W
            wp.s_co, wp.s_ci = 16, cout * 16
        tarr = (WTap * len(wtaps))()
        for i, (v, dx, dy, gv, c0, ci0, nci, off) in enumerate(wtaps):
            tarr[i].view, tarr[i].dx, tarr[i].dy, tarr[i].gview = v, dx, dy, gv
            tarr[i].c0, tarr[i].ci0, tarr[i].nci, tarr[i].off = c0, ci0, nci, off
        wp.n_taps = len(wtaps)
        wp.taps = C.cast(tarr, C.POINTER(WTap))
        wp.cout = cout
        g0, gn = self.grad_slices[sp.w_key]
        wp.dw = self.grad_flat.data_ptr() + 4 * g0
        if sp.name in self.wscratch_slices:
            # scratch destination [tap][co][ci]: tap offset = kernel tap index * cout * cin
            w0, _wn = self.wscratch_slices[sp.name]
            self._wunpack_jobs.append((self.wscratch.data_ptr() + 4 * w0, wp.dw, cout, cin_tot, kk, int(wp.s_co), int(wp.s_ci)))
            self._wunpack_names.append(sp.name)
            for i in range(len(wtaps)):
                tarr[i].off = int(tarr[i].off) * cout * cin_tot
            wp.dw = self.wscratch.data_ptr() + 4 * w0
            wp.s_co, wp.s_ci = cin_tot, 1
        h = C.c_void_p()
        _lib.check(self.L.rnr_wgrad_plan_create(C.byref(wp), self.wgrad_impl, C.byref(h)), 'rnr_wgrad_plan_create')
        st.wgrad_plan = _Plan(h, 'wgrad')

        # ------------------------------ data gradient ------------------------------
        if sp.name == 'in':
            if self.input_grad_range is None:
                return
            r0, r1 = self.input_grad_range
        else:
            r0, r1 = 0, cin_tot
        nci = r1 - r0
        nci_pad = _rup(nci, 16)
        Hi, Wi = sp.H, sp.W
        gC = G.C
        # room for the BatchNorm-backward statistics of the producer layer(s) in this launch's epilogue (_plan_gstats)
        bn_of = {q.dst: q.bn_key for q in self.specs}
        depi = EPI_GSTATS if (self.impl == 1 and self.raw_dt != F32 and os.environ.get('RNR_BN_BWD_FUSED', '0') == '1' and
                              any(bn_of.get(a) is not None for a in sp.src)) else 0
        g_oob = (self.impl == 1 and gC % 64 != 0 and gC % 8 == 0 and os.environ.get('RNR_CONV_OOB', '1') != '0')
        gbk = 64 if g_oob else self._bk_for([gC])
        gK = _rup(gC, 64) if g_oob else gC            # K extent of the gradient operand (chunks past gC are zero-filled by the TMA)
        if sp.kind == 'c3':
            # gxp[ih,iw,ci] = sum_{kh,kw,co} Gp[ih-kh+1, iw-kw+1, co] W[co,ci,kh,kw]   over the padded input plane
            Hp, Wp = Hi + 2, Wi + 2
            st.gx = self._alloc((N, Hp, Wp, nci_pad), gdt_t, zero=True)
            st.gx_fold, st.gx_ld = True, nci_pad
            gch = 1 if gbk == 64 else 0
            tapl = [(1 - kw, 1 - kh, kh * 3 + kw) for kh in range(3) for kw in range(3)]
            ks = self._order_ksteps([(0, dx, dy) for (dx, dy, _) in tapl], gK, gbk, gch)
            tapoffs = [o for (_, _, o) in tapl]
            wm = self._alloc((nci_pad, 9 * gK), gdt_t, zero=True)
            # rows = ci (stride 9), cols = co (stride Cin*9)
            st.wprep_dgrad.append(_WPrep(sp.w_key, r0 * 9, wm, self.grad_dt, nci, nci_pad, cout, gK, 9, 9, cin_tot * 9,
                                         self._tapoff(tapoffs), chunked=gch))
            st.wmat_dgrad = wmat_desc(wm, 9 * gK, self.grad_dt, 9, nci_pad, [tapoffs], r0, r1, ok=bool(gch))
            st.dgrad_plans.append(self._conv_problem(
                [G.padded()], ks, self.grad_dt, gbk, wm, nci_pad, nci_pad, N, Hp, Wp, st.gx, self.grad_dt,
                (Hp * Wp * nci_pad, Wp * nci_pad, nci_pad), (1, 1, 0, 0), depi, None, None, 0, self.impl))
        elif sp.kind == 'c4s2':
            Hp, Wp = Hi + 2, Wi + 2
            st.gx = self._alloc((N, Hp, Wp, nci_pad), gdt_t, zero=True)
            st.gx_fold, st.gx_ld = True, nci_pad
            wm_all = self._alloc((4 * nci_pad, 4 * gK), gdt_t, zero=True)
            probs = []
            dsub_taps = []
            for ph in range(2):
                for pw in range(2):
                    gch = 1 if gbk == 64 else 0
                    tapl = [(1 - b, 1 - a, (2 * a + ph) * 4 + (2 * b + pw)) for a in range(2) for b in range(2)]
                    ks = self._order_ksteps([(0, dx, dy) for (dx, dy, _) in tapl], gK, gbk, gch)
                    tapoffs = [o for (_, _, o) in tapl]
                    dsub_taps.append(tapoffs)
                    sidx = ph * 2 + pw
                    wm = wm_all[sidx * nci_pad:(sidx + 1) * nci_pad]
                    st.wprep_dgrad.append(_WPrep(sp.w_key, r0 * 16, wm, self.grad_dt, nci, nci_pad, cout, gK, 4, 16,
                                                 cin_tot * 16, self._tapoff(tapoffs), chunked=gch))
                    probs.append(self._conv_problem(
                        [G.padded()], ks, self.grad_dt, gbk, wm, nci_pad, nci_pad, N, Hp // 2, Wp // 2, st.gx, self.grad_dt,
                        (Hp * Wp * nci_pad, Wp * nci_pad, nci_pad), (2, 2, ph, pw), depi, None, None, 0, self.impl, defer=True))
            st.wmat_dgrad = wmat_desc(wm_all, 4 * gK, self.grad_dt, 4, nci_pad, dsub_taps, r0, r1, ok=bool(gbk == 64))
            self.keep.append(probs)
            fused = self._fused_plan(probs, self.impl)
            if fused is not None:
                st.dgrad_plans.append(fused)
            else:
                st.dgrad_plans.extend(self._single_plan(pr, self.impl) for pr in probs)
        else:
            # gX[ih,iw,ci] = sum_{kh,kw,co} Gp[2ih+kh, 2iw+kw, co] W[ci,co,kh,kw]
            st.gx = self._alloc((N, Hi, Wi, nci_pad), gdt_t, zero=True)
            st.gx_fold, st.gx_ld = False, nci_pad
            gviews = [G.parity(p, q) for p in range(2) for q in range(2)]
            gch = 1 if gbk == 64 else 0
            # parity-major tap order: the 4 taps reading one parity view of G are adjacent
            tapl = [(p * 2 + q, b, a, (2 * a + p) * 4 + (2 * b + q)) for p in range(2) for q in range(2) for a in range(2) for b in range(2)]
            ks = self._order_ksteps([(v, dx, dy) for (v, dx, dy, _) in tapl], gK, gbk, gch)
            tapoffs = [o for (_, _, _, o) in tapl]
            wm = self._alloc((nci_pad, 16 * gK), gdt_t, zero=True)
            # weight [Cin, Cout, 4,4]: rows = ci (stride Cout*16), cols = co (stride 16)
            st.wprep_dgrad.append(_WPrep(sp.w_key, r0 * cout * 16, wm, self.grad_dt, nci, nci_pad, cout, gK, 16, cout * 16, 16,
                                         self._tapoff(tapoffs), chunked=gch))
            st.wmat_dgrad = wmat_desc(wm, 16 * gK, self.grad_dt, 16, nci_pad, [tapoffs], r0, r1, ok=bool(gch))
            st.dgrad_plans.append(self._conv_problem(
                gviews, ks, self.grad_dt, gbk, wm, nci_pad, nci_pad, N, Hi, Wi, st.gx, self.grad_dt,
                (Hi * Wi * nci_pad, Wi * nci_pad, nci_pad), (1, 1, 0, 0), depi, None, None, 0, self.impl))

    # ------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _mark(self, kind=None, name=None, start=None):
        """CUDA-event bracket around a group of launches on the current stream (only when ``self.timing`` is a list)."""
        if self.timing is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if start is not None:
            self.timing.append((kind, name, start, ev))
        return ev

    @staticmethod
    def layer_flops(sp, N):
        """2*MAC of one live layer (forward); dgrad and wgrad each cost the same."""
        cin = sum(sp.cin)
        if sp.kind == 'ct':
            return 2.0 * N * sp.H * sp.W * cin * sp.cout * 16
        k = 3 if sp.kind == 'c3' else 4
        return 2.0 * N * sp.Ho * sp.Wo * cin * sp.cout * k * k

    def _set_bn(self, st, training, momentum, eps, probe=False):
        """(Re)configure the fused BatchNorm finalize of one layer's forward plan; returns False when the plan cannot host it."""
        sp = st.spec
        N = self.N
        rm = rv = nbt = None
        if training:
            rm, rv = self.buffers[sp.bn_key + '.running_mean'], self.buffers[sp.bn_key + '.running_var']
            nbt = self.buffers.get(sp.bn_key + '.num_batches_tracked')
            if nbt is not None and not (nbt.is_cuda and nbt.dtype == torch.int64):
                nbt = None
        ptr = lambda t: t.data_ptr() if t is not None else None
        rc = self.L.rnr_conv_plan_set_bn(st.fwd_plans[0].h, self.params[sp.bn_key + '.weight'].data_ptr(),
                                         self.params[sp.bn_key + '.bias'].data_ptr(), float(N * sp.Ho * sp.Wo), eps, momentum,
                                         st.mean.data_ptr(), st.invstd.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(),
                                         ptr(rm), ptr(rv), ptr(nbt), st.bn_ticket.data_ptr(), 1 if training else 0)
        if rc != 0 and not probe:
            _lib.check(rc, 'rnr_conv_plan_set_bn(%s)' % sp.name)
        return rc == 0

    def counts_batches_in_kernel(self, bn_key):
        """True when num_batches_tracked of this BatchNorm layer is incremented by the conv kernel's fused finalize (the
        host-side nn.Module bookkeeping must then skip it)."""
        for sp in self.specs:
            if sp.bn_key == bn_key:
                nbt = self.buffers.get(bn_key + '.num_batches_tracked')
                return bool(self.layers[sp.name].bn_fused and nbt is not None and nbt.is_cuda and nbt.dtype == torch.int64)
        return False

    def _wprep(self, items: List[_WPrep], s):
        for w in items:
            if w.ld:
                raise NotImplementedError('per-matrix weight prep has no row pitch: use the batched plan')
            src = self.params[w.src_key]
            _lib.check(self.L.rnr_weight_prep(src.data_ptr() + 4 * w.src_off, w.dst.data_ptr(), w.dtype, w.nr, w.nr_pad,
                                              w.nc, w.cpad, w.ntaps, w.s_r, w.s_c, w.tapoff.data_ptr(), w.chunked, s), 'rnr_weight_prep')
            self.gpu_launches += 1

    def prepare_weights(self, backward=False):
        """fp32 parameters -> 16-bit GEMM matrices of every layer (forward, and data-gradient when ``backward``): one launch."""
        plan = self.wprep_all_plan if backward else self.wprep_fwd_plan
        _lib.check(self.L.rnr_wprep_run(plan.h, self._stream()), 'rnr_wprep_run')
        self.gpu_launches += 1

    def prepare_weights_split(self, part):
        """'fwd' or 'dgrad' half of prepare_weights(backward=True) as its own launch: the fused step runs the data-gradient
        half on a side stream underneath the forward pass."""
        plan = self.wprep_fwd_plan if part == 'fwd' else self.wprep_dgrad_plan
        if plan is None:
            return
        _lib.check(self.L.rnr_wprep_run(plan.h, self._stream()), 'rnr_wprep_run')
        self.gpu_launches += 1

    def wprep_plan_for(self, layer_names, backward=True, skip_fwd=(), skip_dgrad=()):
        """Batched weight-preparation plan (one launch) for the forward (+ data-gradient) matrices of a subset of layers: the
        fused optimiser refreshes each group of layers right after its Adam update.  ``skip_fwd`` / ``skip_dgrad``: layers whose
        matrices the optimiser pass itself writes."""
        items = []
        for sp in self.specs:
            if sp.name in layer_names:
                st = self.layers[sp.name]
                if sp.name not in skip_fwd:
                    items += st.wprep_fwd
                if backward and sp.name not in skip_dgrad:
                    items += st.wprep_dgrad
        return self._wprep_plan(items) if items else None

    def run_wprep_plan(self, plan):
        if plan is not None:
            _lib.check(self.L.rnr_wprep_run(plan.h, self._stream()), 'rnr_wprep_run')
            self.gpu_launches += 1

    def wgrad_scratch_jobs(self, layer_names=None):
        """[(layer name, weight key, scratch ptr, cout, cin, ntaps, s_co, s_ci)] of the layers whose weight gradient is
        accumulated in GEMM order (all of them on the tensor-core path) -- the job table of the fused un-transpose + Adam pass."""
        out = []
        for name, (src, _dst, cout, cin, kk, s_co, s_ci) in zip(self._wunpack_names, self._wunpack_jobs):
            if layer_names is None or name in layer_names:
                out.append((name, self.layers[name].spec.w_key, src, cout, cin, kk, s_co, s_ci))
        return out

    def prepare_weights_per_layer(self, backward=False):
        """Same result through the single-matrix entry point (kept as the cross-check of the batched kernel)."""
        s = self._stream()
        for sp in self.specs:
            st = self.layers[sp.name]
            self._wprep(st.wprep_fwd, s)
            if backward:
                self._wprep(st.wprep_dgrad, s)

    def set_input_nchw(self, x: torch.Tensor):
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        N, Cc, H, W = x.shape
        assert (N, Cc, H, W) == (self.N, self.in_channels, self.H, self.W), 'engine built for a different input shape'
        _lib.check(self.L.rnr_pack_nchw_to_act(x.data_ptr(), self.acts['input'].ptr,
                                               self.acts_w['input'].ptr if self.dual else None, N, Cc, self.in_cpad, H, W, self._stream()),
                   'rnr_pack_nchw_to_act')
        self.gpu_launches += 1

    def forward(self, training: bool, drop_masks: Optional[Dict[str, torch.Tensor]] = None, momentum: float = 0.1,
                eps: float = 1e-5, weights_ready: bool = False, need_backward_prep: Optional[bool] = None):
        """Runs all layers on the current stream.  ``drop_masks[layer name]`` = [N, C] fp32 scale tensor
        (0 or 1/(1-p)) or None for no dropout.  Returns the final fp32 NHWC tensor [N,H,W,ld] (tanh applied)."""
        L, s = self.L, self._stream()
        if not weights_ready:
            self.prepare_weights(backward=self.need_backward)
        self.bn_training = bool(training)
        mode = (bool(training), float(momentum), float(eps))
        if mode != self._bn_mode:
            for sp in self.specs:
                if self.layers[sp.name].bn_fused:
                    self._set_bn(self.layers[sp.name], *mode)
            self._bn_mode = mode
        for sp in self.specs:
            st = self.layers[sp.name]
            t0 = self._mark()
            for pl in st.fwd_plans:
                _lib.check(L.rnr_conv_run(pl.h, s), 'rnr_conv_run(%s)' % sp.name)
                self.gpu_launches += 1
            self._mark('fwd', sp.name, t0)
            if sp.dst == 'out':
                continue
            N, Ho, Wo, Cc = self.N, sp.Ho, sp.Wo, sp.cout
            if sp.bn_key is not None and not training:
                # nn.BatchNorm2d in eval(): normalise with the RUNNING statistics (scale = gamma / sqrt(running_var + eps),
                # shift = beta - running_mean * scale); the batch sums of the conv epilogue are ignored and nothing is updated
                rm, rv = self.buffers[sp.bn_key + '.running_mean'], self.buffers[sp.bn_key + '.running_var']
                with torch.no_grad():
                    torch.add(rv, eps, out=st.invstd).rsqrt_()
                    st.mean.copy_(rm)
                    torch.mul(self.params[sp.bn_key + '.weight'], st.invstd, out=st.scale)
                    torch.addcmul(self.params[sp.bn_key + '.bias'], st.mean, st.scale, value=-1.0, out=st.shift)
                    if st.fwd_totals is not None:
                        st.fwd_totals.zero_()    # (the conv launch added its batch sums; nobody consumes them in eval mode)
                shift = st.shift
            elif sp.bn_key is not None and st.fwd_totals is not None:
                # the conv launch left its batch sums in fp64 totals: the activation pass derives mean / invstd itself
                rm, rv = self.buffers[sp.bn_key + '.running_mean'], self.buffers[sp.bn_key + '.running_var']
                st.drop = drop_masks.get(sp.name) if (drop_masks and sp.drop) else None
                _lib.check(L.rnr_bn_act_fwd_tot(st.raw.data_ptr(), self.raw_dt, st.fwd_totals.data_ptr(), st.fwd_ticket.data_ptr(),
                                                float(N * Ho * Wo), self.params[sp.bn_key + '.weight'].data_ptr(),
                                                self.params[sp.bn_key + '.bias'].data_ptr(), eps, st.mean.data_ptr(),
                                                st.invstd.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(),
                                                rm.data_ptr() if rm is not None else None, rv.data_ptr() if rv is not None else None,
                                                momentum, st.drop.data_ptr() if st.drop is not None else None, sp.slope,
                                                self.acts[sp.dst].ptr, self.acts_w[sp.dst].ptr if self.dual else None,
                                                N, Ho, Wo, Cc, s), 'rnr_bn_act_fwd_tot')
                self.gpu_launches += 1
                continue
            elif sp.bn_key is not None and st.bn_fused:
                shift = st.shift                 # mean / invstd / scale / shift were written by the conv kernel's last CTA
            elif sp.bn_key is not None:
                rm, rv = self.buffers[sp.bn_key + '.running_mean'], self.buffers[sp.bn_key + '.running_var']
                _lib.check(L.rnr_bn_finalize(st.stats.data_ptr(), st.n_stat_tiles, Cc, Cc, float(N * Ho * Wo),
                                             self.params[sp.bn_key + '.weight'].data_ptr(),
                                             self.params[sp.bn_key + '.bias'].data_ptr(), eps,
                                             st.mean.data_ptr(), st.invstd.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(),
                                             rm.data_ptr() if rm is not None else None,
                                             rv.data_ptr() if rv is not None else None, momentum, s), 'rnr_bn_finalize')
                self.gpu_launches += 1
                shift = st.shift
            else:
                shift = self.params[sp.b_key]
            st.drop = drop_masks.get(sp.name) if (drop_masks and sp.drop) else None
            _lib.check(L.rnr_bn_act_fwd(st.raw.data_ptr(), self.raw_dt, st.scale.data_ptr(), shift.data_ptr(),
                                        st.drop.data_ptr() if st.drop is not None else None, sp.slope,
                                        self.acts[sp.dst].ptr, self.acts_w[sp.dst].ptr if self.dual else None,
                                        N, Ho, Wo, Cc, s), 'rnr_bn_act_fwd')
            self.gpu_launches += 1
        return self.layers['out'].raw

    def output_nchw(self) -> torch.Tensor:
        sp = self.specs[-1]
        out = torch.empty((self.N, sp.cout, sp.Ho, sp.Wo), dtype=torch.float32, device=self.device)
        _lib.check(self.L.rnr_unpack_nhwc_to_nchw(self.layers['out'].raw.data_ptr(), out.data_ptr(), self.N, sp.cout, self.out_ld,
                                                  sp.Ho, sp.Wo, self._stream()), 'rnr_unpack_nhwc_to_nchw')
        self.gpu_launches += 1
        return out

    def gate_masks(self) -> Dict[str, torch.Tensor]:
        """(pre-activation > 0) per live layer as NCHW bool tensors -- test hook: lets the oracle differentiate
        through the same ReLU/LeakyReLU gates as the engine."""
        out = {}
        for sp in self.specs[:-1]:
            st = self.layers[sp.name]
            shift = st.shift if sp.bn_key is not None else self.params[sp.b_key]
            z = st.raw.float() * st.scale + shift
            out[sp.name] = (z > 0).permute(0, 3, 1, 2).contiguous()
        return out

    def grad_view(self, key) -> torch.Tensor:
        o, n = self.grad_slices[key]
        return self.grad_flat[o:o + n].view(self.params[key].shape)

    def _gsrcs_for(self, act_name):
        """Gradient sources of an activation tensor = dgrad results of its consumers."""
        out = []
        for (lname, si) in self.consumers.get(act_name, []):
            st = self.layers[lname]
            if st.gx is None:
                continue
            g = GSrc()
            g.ptr = st.gx.data_ptr()
            g.dtype = self.grad_dt
            g.fold = 1 if st.gx_fold else 0
            g.ld = st.gx_ld
            g.c0 = sum(st.spec.cin[:si])
            out.append(g)
        return out

    def zero_grads(self):
        """Zero the accumulation targets of the backward pass (flat gradient buffer + weight-gradient scratch)."""
        self.grad_flat.zero_()
        if self.wscratch is not None:
            self.wscratch.zero_()

    def backward_from_nchw(self, grad_out: torch.Tensor):
        """grad_out: d loss / d (tanh output), NCHW fp32.  Fills self.grad_flat; returns grad wrt the input
        channels ``input_grad_range`` as NCHW fp32 (or None)."""
        assert self.need_backward
        L, s = self.L, self._stream()
        sp = self.specs[-1]
        st = self.layers['out']
        self.zero_grads()
        go = grad_out.contiguous()
        th = st.raw
        if not self.final_tanh:      # linear output: d/d raw = grad * (1 - 0^2)
            if self._zeros_out is None:
                self._zeros_out = torch.zeros_like(st.raw)
            th = self._zeros_out
        _lib.check(L.rnr_tanh_bwd_pack(go.data_ptr(), th.data_ptr(), self.gz['out'].ptr,
                                       self.grad_view(sp.b_key).data_ptr(), self.N, sp.cout, self.out_ld, sp.Ho, sp.Wo, s),
                   'rnr_tanh_bwd_pack')
        self.gpu_launches += 1
        return self._backward_layers()

    def _backward_layers(self, after_layer=None, before_unpack=None, unpack=True, input_grad_add=None):
        """Backward of every layer in reverse order.  ``after_layer(name)`` is called once a layer's weight- and data-gradient
        launches are enqueued, ``before_unpack()`` before the weight gradients leave GEMM order (hooks of the data-parallel
        step: early all-reduce of ``wscratch``)."""
        L, s = self.L, self._stream()
        N = self.N
        if not getattr(self, 'bn_training', True) and any(sp.bn_key for sp in self.specs):
            raise NotImplementedError('backward through BatchNorm in eval() mode (running statistics) is not implemented: the '
                                      'reference scripts always differentiate with BatchNorm in train() mode')
        self._refresh_gstats()
        for sp in reversed(self.specs):
            st = self.layers[sp.name]
            Ho, Wo, Cc = sp.Ho, sp.Wo, sp.cout
            if sp.dst != 'out' and sp.name in self.gstat_layers:
                # the consumers' data-gradient launches already accumulated sum(gg), sum(gg * (raw - mean)): one apply pass
                srcs = self._gsrcs_for(sp.dst)
                arr = (GSrc * len(srcs))(*srcs)
                _lib.check(L.rnr_bn_bwd_apply_src(arr, len(srcs), st.raw.data_ptr(), self.raw_dt, st.scale.data_ptr(), st.shift.data_ptr(),
                                                  st.mean.data_ptr(), st.invstd.data_ptr(), self.params[sp.bn_key + '.weight'].data_ptr(),
                                                  st.drop.data_ptr() if st.drop is not None else None, sp.slope, self.gz[sp.name].ptr,
                                                  st.bwd_totals.data_ptr(), st.ticket.data_ptr(), float(N * Ho * Wo),
                                                  self.grad_view(sp.bn_key + '.weight').data_ptr(),
                                                  self.grad_view(sp.bn_key + '.bias').data_ptr(), N, Ho, Wo, Cc, s), 'rnr_bn_bwd_apply_src')
                self.gpu_launches += 1
            elif sp.dst != 'out':
                srcs = self._gsrcs_for(sp.dst)
                assert 1 <= len(srcs) <= 2, (sp.name, len(srcs))
                arr = (GSrc * len(srcs))(*srcs)
                has_bn = sp.bn_key is not None
                shift = st.shift if has_bn else self.params[sp.b_key]
                if has_bn:
                    dgam = self.grad_view(sp.bn_key + '.weight').data_ptr()
                    dbet = self.grad_view(sp.bn_key + '.bias').data_ptr()
                else:
                    dgam, dbet = None, self.grad_view(sp.b_key).data_ptr()
                gam = self.params[sp.bn_key + '.weight'].data_ptr() if has_bn else None
                # pass 1 (activation' / dropout gate, per-channel sums) with the finalize fused into its last block
                _lib.check(L.rnr_bn_bwd_reduce_fin(arr, len(srcs), st.raw.data_ptr(), self.raw_dt, st.scale.data_ptr(), shift.data_ptr(),
                                                   st.mean.data_ptr(), st.invstd.data_ptr(),
                                                   st.drop.data_ptr() if st.drop is not None else None, sp.slope,
                                                   self.gz[sp.name].ptr, st.bwd_totals.data_ptr(), st.ticket.data_ptr(),
                                                   float(N * Ho * Wo), dgam, dbet, gam, st.coef.data_ptr() if has_bn else None,
                                                   N, Ho, Wo, Cc, s), 'rnr_bn_bwd_reduce_fin')
                self.gpu_launches += 1
                if has_bn:
                    _lib.check(L.rnr_bn_bwd_apply(self.gz[sp.name].ptr, st.raw.data_ptr(), self.raw_dt, st.coef.data_ptr(), N, Ho, Wo, Cc, s),
                               'rnr_bn_bwd_apply')
                    self.gpu_launches += 1
            t0 = self._mark()
            _lib.check(L.rnr_wgrad_run(st.wgrad_plan.h, s), 'rnr_wgrad_run(%s)' % sp.name)
            self.gpu_launches += 1
            t1 = self._mark('wgrad', sp.name, t0)
            for pl in st.dgrad_plans:
                _lib.check(L.rnr_conv_run(pl.h, s), 'rnr_conv_run(dgrad %s)' % sp.name)
                self.gpu_launches += 1
            if st.dgrad_plans:
                self._mark('dgrad', sp.name, t1)
            if after_layer is not None:
                after_layer(sp.name)
        if before_unpack is not None:
            before_unpack()
        # ``unpack=False``: the caller consumes the GEMM-order scratch itself (fused un-transpose + Adam, rnr_adam_run)
        if unpack and self.wunpack_plan is not None:
            _lib.check(L.rnr_wgrad_unpack_run(self.wunpack_plan.h, s), 'rnr_wgrad_unpack_run')
            self.gpu_launches += 1
        st = self.layers['in']
        if st.gx is None:
            return None
        r0, r1 = self.input_grad_range
        if self._gi is None or self._gi.shape[1] != r1 - r0:
            self._gi = torch.empty((N, r1 - r0, self.H, self.W), dtype=torch.float32, device=self.device)
        gi = self._gi if input_grad_add is not None else torch.empty_like(self._gi)
        if input_grad_add is not None:
            # ``input_grad_add`` [N, k, H, W] is added to the first k channels in the same pass (the fused step's albedo gradient)
            _lib.check(L.rnr_fold_to_nchw_add(st.gx.data_ptr(), self.grad_dt, gi.data_ptr(), N, r1 - r0, 0, st.gx_ld, self.H, self.W,
                                              input_grad_add.data_ptr(), int(input_grad_add.shape[1]), s), 'rnr_fold_to_nchw_add')
        else:
            _lib.check(L.rnr_fold_to_nchw(st.gx.data_ptr(), self.grad_dt, gi.data_ptr(), N, r1 - r0, 0, st.gx_ld, self.H, self.W, s),
                       'rnr_fold_to_nchw')
        self.gpu_launches += 1
        return gi

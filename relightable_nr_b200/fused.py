"""The RNR training / rendering step with its per-pixel stages fused around the U-Net (csrc/fused.cu) and the optimiser
fused into the backward pass (csrc/optim.cu).

Same computation as ``RNRPipeline.forward / losses / backward / optimizer.step`` (train_rnr.py:490-623) -- those drive the drop-in
modules one operator at a time, exactly like the reference script, and remain the parity reference for this file -- but:

* ``rnr_head_fwd`` writes the first convolution's operand (fp16 channels-last, reflect halo) directly from the texture
  pyramid and the per-view maps: no [N,108,H,W] fp32 tensor, no permute / cat / pack;
* ``rnr_tail_fwd / rnr_tail_bwd`` read the last convolution's NHWC output and write the data-gradient operand (bf16, zero
  halo), the bias gradient, the albedo and envmap gradients: no [N,26,3,H,W] temporaries;
* the two small losses (lighting L1, albedo mean) are four kernels on a side stream (csrc/smallloss.cu), the loss scalar is
  combined on the device, Dropout2d masks come from a counter-based generator: no ATen kernel is left in the step;
* **optimiser** (torch.optim.Adam semantics, train_rnr.py:376): the weight gradients stay in the GEMM order the tcgen05
  kernels accumulate them in; ``rnr_adam_run`` un-transposes, applies Adam to the fp32 master weights and re-zeroes the
  scratch in ONE pass, and the 16-bit GEMM matrices of the next step are re-derived right behind it.  The layers that finish
  first in backward order (95 % of the parameters) are updated on a second stream underneath the backward pass of the
  full-resolution layers -- in a data-parallel run right behind their all-reduce.  Small tensors (biases, BatchNorm affine,
  texture levels, SH coefficients) are one multi-tensor launch that also advances the device-resident step counter.

No autograd graph is built: the backward is the explicit kernel sequence below.  With ``step_optimizer=False`` (parity tests)
the gradients are materialised in parameter layout instead (``param.grad`` views of the engine's flat buffer) and nothing is
updated.
"""
import ctypes as C
import os

import torch

from . import _lib, ops
from ._lib import AdamJob, AdamWJob

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
u64 = C.c_ulonglong
_pp = C.POINTER(C.c_void_p)
_ip = C.POINTER(C.c_int)
_lib.register_sigs({
    "rnr_head_fwd": [_pp, _ip, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, i32, i32, vp],
    "rnr_tail_fwd": [vp, i32, vp, vp, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "rnr_tail_bwd": [vp, i32, vp, vp, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, f32, f32, vp, i32, vp, vp, vp, vp],
    "rnr_adam_plan_create": [C.POINTER(AdamWJob), i32, C.POINTER(AdamJob), i32, C.POINTER(vp)],
    "rnr_adam_run": [vp, vp, f32, f32, f32, f32, f32, i32, i32, vp],
    "rnr_loss_combine": [vp, f64, f64, f64, vp, i32, vp, i32, vp],
    "rnr_dropout_masks": [vp, i32, f32, u64, vp, vp],
    "rnr_lighting_l1": [vp, vp, vp, vp, i32, i32, f32, f32, vp, vp, vp],
    "rnr_albedo_mean_loss": [vp, vp, i64, f32, vp, vp, vp, vp],
    "rnr_sh_reconstruct_ld": [vp, vp, vp, i64, i32, i32, i32, vp],
    "rnr_sh_project_ld": [vp, vp, vp, i64, i32, i32, i32, f32, i32, vp],
})


def _s():
    return torch.cuda.current_stream().cuda_stream


def _cf(t):
    """fp32 contiguous view of a per-view map (a no-op for the maps ViewDataset / synthetic_view produce)."""
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _AdamPlan:
    def __init__(self, L, wjobs, jobs):
        self.L = L
        wa = (AdamWJob * max(len(wjobs), 1))()
        for d, (scratch, p, m, v, cout, cin, ntaps, s_co, s_ci, fwd, dgrad) in zip(wa, wjobs):
            d.scratch, d.p, d.m, d.v, d.gdst = scratch, p, m, v, None
            d.cout, d.cin, d.ntaps, d.s_co, d.s_ci = cout, cin, ntaps, s_co, s_ci
            for dst, desc in ((d.fwd, fwd), (d.dgrad, dgrad)):
                dst.base = None
                if desc is not None:      # the pass also writes next step's 16-bit GEMM matrix of this family (engine/unet.py wmat_desc)
                    dst.base, dst.ld, dst.dtype = desc['base'].data_ptr(), desc['ld'], desc['dtype']
                    dst.ntaps, dst.sub_rows, dst.r0, dst.r1 = desc['ntaps'], desc['sub_rows'], desc['r0'], desc['r1']
                    for t in range(16):
                        dst.inv[t] = desc['inv'][t]
        ja = (AdamJob * max(len(jobs), 1))()
        for d, (p, g, m, v, n) in zip(ja, jobs):
            d.p, d.g, d.m, d.v, d.n = p, g, m, v, n
        self.h = C.c_void_p()
        _lib.check(L.rnr_adam_plan_create(wa, len(wjobs), ja, len(jobs), C.byref(self.h)), 'rnr_adam_plan_create')
        self.launches = (1 if wjobs else 0) + (1 if jobs else 0)

    def __del__(self):
        try:
            if self.h:
                self.L.rnr_adam_plan_destroy(self.h)
        except Exception:
            pass


class FusedRNRStep:
    """Fused execution of one RNRPipeline iteration.  Built lazily per (N, H, W); owns scratch buffers and, once it has taken an
    optimiser step, the Adam state of every parameter of the step (``optimizer_state()``)."""

    CROP = 5

    def __init__(self, pipe):
        self.pipe = pipe
        self.dev = pipe.device
        self.L = _lib.lib()
        self._shape = None
        self.side = torch.cuda.Stream(device=self.dev)
        self.grad_hook = None          # callable(list of gradient tensors) between backward and the optimiser (step_optimizer=False / legacy)
        #: data parallel: callable(tensor) that SUMS one gradient buffer over the ranks in place (enqueued on the current stream);
        #: the 1/world factor is applied inside the Adam kernels (``world``).  The weight gradients of the layers that finish
        #: first in backward order go out, still in GEMM order, while the full-resolution layers are being differentiated.
        self.allreduce_sum = None
        self.world = 1
        #: legacy hook: callable(tensor) that all-reduces AND scales (mean); used when ``allreduce_sum`` is None
        self.allreduce = None
        self.comm = torch.cuda.Stream(device=self.dev)     # early all-reduce + early optimiser group
        self.early_layer = 'b3.down1'  # last layer (in backward order) of the early group
        self._opt = None
        self._wver = None
        self._grads_clean = False
        self._sums_clean = False

    # ------------------------------------------------------------------------------------------------------------------
    def _setup(self, N, H, W, need_backward):
        key = (N, H, W, need_backward)
        if self._shape == key:
            return
        p, dev = self.pipe, self.dev
        tm = p.texture_mapper
        self.C = int(tm.textures[0].shape[-1])
        self.Rs, self.Rd = int(p.ray_sampler.num_ray), int(p.ray_sampler_diffuse.num_ray)
        self.R = self.Rs + self.Rd
        cin = 3 * self.R + 6 + self.C
        runner = p.render_net.net._runner
        rng = (3 * self.R + 6, cin) if need_backward else None
        self.eng = runner.engine_for(dev, N, H, W, need_backward, True, rng)
        self.runner = runner
        f32k = dict(dtype=torch.float32, device=dev)
        self.rays_uv = torch.empty((N, H, W, 2, self.R), **f32k)
        self.albedo = torch.zeros((N, H, W, 8), **f32k)
        self.aux = torch.empty((N, H, W, 12), **f32k)
        self.final = torch.empty((N, 3, H, W), **f32k)
        # device accumulators: [0:4] tail kernels (chrom numerator / denominator, L1 sum, -), [4] lighting loss, [5] albedo loss,
        # [6:14] albedo-mean counts and channel sums
        self.sums = torch.zeros(16, dtype=torch.float64, device=dev)
        self.loss_out = torch.zeros(1, **f32k)
        lm = p.lighting_model
        self.Hl, self.Wl = int(lm.lp_recon_h), int(lm.lp_recon_w)
        self.B = int(lm.basis_val.shape[1])
        self.lp4 = torch.zeros((self.Hl * self.Wl, 4), **f32k)          # envmap texels (r, g, b, 0)
        self._unet_mod = runner._unet()
        # Dropout2d masks: one buffer [N, sum of channels], views per layer; counter-based generator state
        eng = self.eng
        self._drop_names = [sp.name for sp in eng.specs if sp.drop and sp.dst != 'out']
        chans = [eng.layers[n].spec.cout for n in self._drop_names]
        self._drop_buf = torch.zeros((sum(chans) * N,), **f32k)
        self._drop_views, o = {}, 0
        for n, c in zip(self._drop_names, chans):
            self._drop_views[n] = self._drop_buf[o:o + N * c].view(N, c)
            o += N * c
        self._drop_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        self._drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if need_backward:
            self.g_alb = torch.empty((N, 6, H, W), **f32k)
            self.g_lp4 = torch.zeros((self.Hl * self.Wl, 4), **f32k)
            # parameter gradients: persistent views (U-Net: into the engine's flat buffer; textures + SH coefficients: one flat
            # buffer, so that a data-parallel step needs ONE collective for them)
            seen = set()
            for k, prm in runner._unet().named_parameters(remove_duplicate=False):
                if k in eng.grad_slices and id(prm) not in seen:
                    prm.grad = eng.grad_view(k)
                    seen.add(id(prm))
            sizes = [t.numel() for t in tm.textures] + [lm.coeff.numel()]
            offs = [0]
            for n in sizes:
                offs.append(offs[-1] + (n + 3) // 4 * 4)
            self.aux_flat = torch.zeros(offs[-1], **f32k)
            self.tex_grads = [self.aux_flat[offs[i]:offs[i] + sizes[i]].view(t.shape) for i, t in enumerate(tm.textures)]
            for t, g in zip(tm.textures, self.tex_grads):
                t.grad = g
            self.coeff_grad = self.aux_flat[offs[-2]:offs[-2] + sizes[-1]].view(lm.coeff.shape)
            lm.coeff.grad = self.coeff_grad
            # small losses
            S = int(lm.basis_val.shape[0])
            self.S = S
            self.sgn = torch.zeros((S, 3), **f32k)
            self.mask_u8 = p.l_samples_init_mask.to(torch.uint8).contiguous()
            n_cov = float(self.mask_u8.sum().item())
            n_unc = float(S - n_cov)
            self.w_cov = float(p.w['lighting']) / n_cov if n_cov > 0 else 0.0
            self.w_unc = float(p.w['lighting_uncovered']) / n_unc if n_unc > 0 else 0.0
            S0 = int(tm.textures[0].shape[1])
            self.tex6 = torch.empty((S0 * S0, 6), **f32k)
            self.gout6 = torch.empty((S0 * S0, 6), **f32k)
            self._gtex_ptrs = (C.c_void_p * len(tm.textures))(*[g.data_ptr() for g in self.tex_grads])
            self._opt = None
            self._grads_clean = False
        self._tex_ptrs = (C.c_void_p * len(tm.textures))(*[t.data_ptr() for t in tm.textures])
        self._tex_sizes = (C.c_int * len(tm.textures))(*[int(t.shape[1]) for t in tm.textures])
        self._tex_ids = [t.data_ptr() for t in tm.textures]
        self._shape = key

    def grad_tensors(self):
        """Every gradient buffer of the step (what a data-parallel all-reduce must average when the optimiser is not fused)."""
        return [self.eng.grad_flat, self.aux_flat]

    # ------------------------------------------------------------------------------------------------------------------
    # optimiser state
    # ------------------------------------------------------------------------------------------------------------------
    def _build_optimizer(self):
        p, eng, dev = self.pipe, self.eng, self.dev
        grp = p.optimizer.param_groups[0]
        self.lr, (self.beta1, self.beta2), self.eps = float(grp['lr']), grp['betas'], float(grp['eps'])
        if grp.get('weight_decay', 0) or grp.get('amsgrad', False):
            raise NotImplementedError('fused optimiser: plain Adam only (weight_decay = 0, amsgrad = False), as train_rnr.py:376 uses it')
        tm, lm = p.texture_mapper, p.lighting_model
        opt = {'m': {}, 'v': {}, 'step': torch.zeros(1, dtype=torch.float32, device=dev)}
        zl = lambda t: torch.zeros_like(t, memory_format=torch.contiguous_format)
        early, late = self._groups()
        wjobs = {True: [], False: []}
        fuse_w = os.environ.get('RNR_ADAM_WRITES_GEMM', '1') != '0'
        skip_fwd, skip_dgrad = set(), set()
        for name, w_key, src, cout, cin, kk, s_co, s_ci in eng.wgrad_scratch_jobs():
            prm = eng.params[w_key]
            opt['m'][w_key], opt['v'][w_key] = zl(prm), zl(prm)
            st = eng.layers[name]
            fwd = st.wmat_fwd if fuse_w else None
            dgr = st.wmat_dgrad if fuse_w else None
            if fwd is not None:
                skip_fwd.add(name)
            if dgr is not None or not st.wprep_dgrad:
                skip_dgrad.add(name)
            wjobs[name in early].append((src, prm.data_ptr(), opt['m'][w_key].data_ptr(), opt['v'][w_key].data_ptr(), cout, cin, kk, s_co, s_ci,
                                         fwd, dgr))
        jobs = []
        done_w = {j[1] for j in eng.wgrad_scratch_jobs()}
        for key, (o, n) in eng.grad_slices.items():
            if key in done_w:
                continue
            prm = eng.params[key]
            opt['m'][key], opt['v'][key] = zl(prm), zl(prm)
            jobs.append((prm.data_ptr(), eng.grad_flat.data_ptr() + 4 * o, opt['m'][key].data_ptr(), opt['v'][key].data_ptr(), n))
        for i, (t, g) in enumerate(zip(tm.textures, self.tex_grads)):
            if t.requires_grad:
                k = 'textures.%d' % i
                opt['m'][k], opt['v'][k] = zl(t), zl(t)
                jobs.append((t.data_ptr(), g.data_ptr(), opt['m'][k].data_ptr(), opt['v'][k].data_ptr(), t.numel()))
        if lm.coeff.requires_grad:
            opt['m']['coeff'], opt['v']['coeff'] = zl(lm.coeff), zl(lm.coeff)
            jobs.append((lm.coeff.data_ptr(), self.coeff_grad.data_ptr(), opt['m']['coeff'].data_ptr(), opt['v']['coeff'].data_ptr(),
                         lm.coeff.numel()))
        opt['plan_early'] = _AdamPlan(self.L, wjobs[True], []) if wjobs[True] else None
        opt['plan_late'] = _AdamPlan(self.L, wjobs[False], []) if wjobs[False] else None
        opt['plan_small'] = _AdamPlan(self.L, [], jobs)
        # weight preparation left over after the optimiser pass: only the matrix families whose layout that pass does not write
        opt['wprep_early'] = eng.wprep_plan_for(early, skip_fwd=skip_fwd, skip_dgrad=skip_dgrad)
        opt['wprep_late'] = eng.wprep_plan_for(late, skip_fwd=skip_fwd, skip_dgrad=skip_dgrad)
        opt['gemm_in_adam'] = (sorted(skip_fwd), sorted(skip_dgrad))
        self._opt = opt

    def _groups(self):
        """(early, late) layer-name sets: ``early`` = every layer from the last one down to ``early_layer`` in forward (spec)
        order -- the first to be differentiated; their weight gradients occupy the tail ``wscratch[w0:]``."""
        names = [sp.name for sp in self.eng.specs]
        if self.early_layer in names:
            i = names.index(self.early_layer)
            return set(names[i:]), set(names[:i])
        return set(names), set()

    def optimizer_state(self):
        """{'step': device float, 'exp_avg': {key: tensor}, 'exp_avg_sq': {key: tensor}} of the fused optimiser (parameter layout;
        keys: U-Net state-dict keys relative to ``Unet``, 'textures.i', 'coeff')."""
        if self._opt is None:
            return None
        return {'step': self._opt['step'], 'exp_avg': self._opt['m'], 'exp_avg_sq': self._opt['v']}

    def _adam(self, plan, advance=False):
        if plan is None:
            return
        _lib.check(self.L.rnr_adam_run(plan.h, self._opt['step'].data_ptr(), self.lr, self.beta1, self.beta2, self.eps,
                                       1.0 / float(self.world), 1, 1 if advance else 0, _s()), 'rnr_adam_run')
        self.eng.gpu_launches += plan.launches

    # ------------------------------------------------------------------------------------------------------------------
    def _head(self, view):
        p, eng = self.pipe, self.eng
        N, H, W = eng.N, eng.H, eng.W
        tm = p.texture_mapper
        if [t.data_ptr() for t in tm.textures] != self._tex_ids:
            raise RuntimeError('texture parameters were re-allocated; rebuild the fused step')
        _lib.check(self.L.rnr_head_fwd(
            C.cast(self._tex_ptrs, _pp), C.cast(self._tex_sizes, _ip), len(tm.textures), self.C,
            _cf(view['uv_map']).data_ptr(), _cf(view['sh_basis_map']).data_ptr(), 6,
            _cf(view['TBN_map']).data_ptr(), _cf(view['view_dir_map_tangent']).data_ptr(), _cf(view['alpha_map']).data_ptr(),
            _cf(view['normal_map']).data_ptr(), _cf(view['view_dir_map']).data_ptr(),
            p.ray_sampler.pivots_dir.data_ptr(), self.Rs, p.ray_sampler_diffuse.pivots_dir.data_ptr(), self.Rd,
            eng.acts['input'].ptr, eng.acts_w['input'].ptr if eng.dual else None, eng.in_cpad,
            self.rays_uv.data_ptr(), self.albedo.data_ptr(), N, H, W, _s()), 'rnr_head_fwd')

    def _tail_fwd(self, view, final=None):
        eng = self.eng
        raw = eng.layers['out'].raw
        final = self.final if final is None else final
        _lib.check(self.L.rnr_tail_fwd(raw.data_ptr(), eng.out_ld, self.rays_uv.data_ptr(), self.albedo.data_ptr(), self.lp4.data_ptr(),
                                       self.Hl, self.Wl, _cf(view['alpha_map']).data_ptr(), _cf(view['img_gt']).data_ptr(), self.Rs, self.Rd,
                                       eng.N, eng.H, eng.W, self.CROP, final.data_ptr(), self.aux.data_ptr(),
                                       self.sums.data_ptr(), _s()), 'rnr_tail_fwd')

    def _coeff_row_ptr(self, tensor):
        lm = self.pipe.lighting_model
        return tensor.data_ptr() + 4 * int(self.pipe.lighting_idx) * int(lm.coeff.shape[1]) * int(lm.coeff.shape[2])

    def _envmap(self):
        """LightingSH.reconstruct_lp (network.py:622-627) of the current coefficients into [Hl*Wl, 4] texels (r, g, b, 0): the tail
        kernels fetch one 16-byte texel per bilinear tap."""
        lm = self.pipe.lighting_model
        _lib.check(self.L.rnr_sh_reconstruct_ld(lm.basis_val_recon.data_ptr(), self._coeff_row_ptr(lm.coeff), self.lp4.data_ptr(),
                                                self.Hl * self.Wl, self.B, 3, 4, _s()), 'rnr_sh_reconstruct_ld')
        return self.lp4

    def _step_state(self):
        """nn.Module bookkeeping of one forward, without ATen launches: (BatchNorm in train mode?, Dropout2d masks or None)."""
        unet, eng = self._unet_mod, self.eng
        bn_mod, drop_mod = unet.in_layer[1], unet.in_layer[3]
        training_bn = bool(bn_mod.training)
        drop = None
        if drop_mod.training and drop_mod.p > 0:
            _lib.check(self.L.rnr_dropout_masks(self._drop_buf.data_ptr(), self._drop_buf.numel(), float(drop_mod.p), self._drop_seed,
                                                self._drop_ctr.data_ptr(), _s()), 'rnr_dropout_masks')
            drop = self._drop_views
        if training_bn:
            nbt = [b for k, b in unet.named_buffers() if k.endswith('num_batches_tracked') and '.fuse.' not in k
                   and not eng.counts_batches_in_kernel(k[:-len('.num_batches_tracked')])]
            if nbt:
                torch._foreach_add_(nbt, 1)
        return training_bn, drop

    # ------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def render(self, view):
        """test_rnr.py:335-371 for one view: returns the final image [N,3,H,W] (valid until the next call)."""
        N, H, W = view['alpha_map'].shape
        if self._shape != (N, H, W, True):          # (a training set-up of the same shape renders too: keep it and its Adam state)
            self._setup(N, H, W, False)
        eng = self.eng
        training_bn, drop = self._step_state()
        self._envmap()
        self._head(view)
        eng.forward(training=training_bn, drop_masks=drop)
        self.sums.zero_()
        self._tail_fwd(view)
        self.sums.zero_()                           # leave the accumulators clean for a (possibly graph-captured) training step
        return self.final

    @torch.no_grad()
    def render_relight(self, view, coeffs):
        """Relighting (test_rnr.py:335-371 over several lightings): ONE texture / ray / U-Net pass per view, then one RayRenderer pass
        per SH coefficient set ``coeffs[k]`` [num_basis, 3] (contiguous CUDA fp32).  Returns a list of images [N,3,H,W] (valid until
        the next call)."""
        N, H, W = view['alpha_map'].shape
        if self._shape != (N, H, W, True):
            self._setup(N, H, W, False)
        eng = self.eng
        training_bn, drop = self._step_state()
        self._head(view)
        eng.forward(training=training_bn, drop_masks=drop)
        if getattr(self, '_relit', None) is None or len(self._relit) != len(coeffs) or self._relit[0].shape != self.final.shape:
            self._relit = [torch.empty_like(self.final) for _ in coeffs]
        lm = self.pipe.lighting_model
        for c, out in zip(coeffs, self._relit):
            if not (c.is_cuda and c.dtype == torch.float32 and c.is_contiguous() and tuple(c.shape) == (self.B, 3)):
                raise ValueError('render_relight: coefficient sets must be contiguous CUDA fp32 [%d, 3] tensors' % self.B)
            _lib.check(self.L.rnr_sh_reconstruct_ld(lm.basis_val_recon.data_ptr(), c.data_ptr(), self.lp4.data_ptr(), self.Hl * self.Wl, self.B,
                                                    3, 4, _s()), 'rnr_sh_reconstruct_ld')
            self.sums.zero_()
            self._tail_fwd(view, out)
        self.sums.zero_()
        return self._relit

    def _small_losses(self):
        """lighting L1 (train_rnr.py:571-579) + albedo-mean loss (train_rnr.py:596-607): values into ``sums[4:6]``, gradients
        accumulated into the (pre-zeroed) texture / coefficient gradient buffers.  Five kernels, no autograd."""
        p, L = self.pipe, self.L
        lm, tm = p.lighting_model, p.texture_mapper
        if lm.coeff.requires_grad:
            _lib.check(L.rnr_lighting_l1(lm.basis_val.data_ptr(), self._coeff_row_ptr(lm.coeff), p.l_samples_init.data_ptr(),
                                         self.mask_u8.data_ptr(), self.S, self.B, self.w_cov, self.w_unc, self.sgn.data_ptr(),
                                         self.sums.data_ptr() + 8 * 4, _s()), 'rnr_lighting_l1')
            _lib.check(L.rnr_sh_project(lm.basis_val.data_ptr(), self.sgn.data_ptr(), self._coeff_row_ptr(self.coeff_grad), self.S, self.B,
                                        3, 1, 1.0, _s()), 'rnr_sh_project')
        nl = len(tm.textures)
        _lib.check(L.rnr_flatten_mipmap(C.cast(self._tex_ptrs, _pp), None, C.cast(self._tex_sizes, _ip), nl, self.C, 0, 6,
                                        self.tex6.data_ptr(), None, 0, _s()), 'rnr_flatten_mipmap')
        _lib.check(L.rnr_albedo_mean_loss(self.tex6.data_ptr(), tm.tex_flatten_mipmap_init.data_ptr(), self.tex6.shape[0], float(p.w['alb']),
                                          self.sums.data_ptr() + 8 * 6, self.gout6.data_ptr(), self.sums.data_ptr() + 8 * 5, _s()),
                   'rnr_albedo_mean_loss')
        if any(t.requires_grad for t in tm.textures):
            _lib.check(L.rnr_flatten_mipmap(C.cast(self._tex_ptrs, _pp), C.cast(self._gtex_ptrs, _pp), C.cast(self._tex_sizes, _ip), nl, self.C,
                                            0, 6, None, self.gout6.data_ptr(), 1, _s()), 'rnr_flatten_mipmap(bwd)')

    def _weights_current(self):
        """The 16-bit GEMM matrices match the fp32 parameters iff nobody but this class touched them since the last refresh:
        torch bumps a tensor's version on every in-place write (optimizer.step, load_state_dict, ...); our kernels do not."""
        ver = tuple(p._version for p in self.eng.params.values())
        return self._wver is not None and ver == self._wver

    def _mark_weights_current(self):
        self._wver = tuple(p._version for p in self.eng.params.values())

    def train_step(self, view, step_optimizer=True):
        """One iteration (train_rnr.py:490-623): returns (loss, final image)."""
        p = self.pipe
        N, H, W = view['alpha_map'].shape
        self._setup(N, H, W, True)
        eng, L = self.eng, self.L
        main = torch.cuda.current_stream(self.dev)
        side = self.side
        training_bn, drop = self._step_state()
        fused_opt = bool(step_optimizer) and eng.wscratch is not None and len(eng.wscratch_slices) == len(eng.specs) \
            and self.grad_hook is None
        if fused_opt and self._opt is None:
            self._build_optimizer()
        ar_sum = self.allreduce_sum
        gscale_in_adam = fused_opt and ar_sum is not None

        # ---- side stream: weight matrices (only when somebody else changed the parameters), zero-fills (only when the last
        # ---- step did not clean up behind itself), the two small losses ----
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ev_w = None
            if not self._weights_current():
                eng.prepare_weights(backward=True)
                self._mark_weights_current()
                ev_w = torch.cuda.Event()
                ev_w.record(side)
            if not self._grads_clean:
                eng.zero_grads()
                self.g_lp4.zero_()
                self.aux_flat.zero_()
            if not self._sums_clean:
                self.sums.zero_()
            ev_z = torch.cuda.Event()
            ev_z.record(side)
            self._small_losses()
        # ---- main stream: envmap, head, U-Net forward, tail ----
        self._envmap()
        self._head(view)
        if ev_w is not None:
            main.wait_event(ev_w)
        eng.forward(training=training_bn, drop_masks=drop, weights_ready=True)
        main.wait_event(ev_z)
        self._tail_fwd(view)
        sp = eng.specs[-1]
        raw = eng.layers['out'].raw
        _lib.check(L.rnr_tail_bwd(raw.data_ptr(), eng.out_ld, self.rays_uv.data_ptr(), self.albedo.data_ptr(), self.lp4.data_ptr(),
                                  self.Hl, self.Wl, _cf(view['alpha_map']).data_ptr(), _cf(view['img_gt']).data_ptr(), self.Rs, self.Rd,
                                  N, H, W, self.CROP, self.aux.data_ptr(), self.sums.data_ptr(), 1.0, float(p.w['rays_lt_chrom']),
                                  eng.gz['out'].ptr, eng.out_ld, eng.grad_view(sp.b_key).data_ptr(), self.g_alb.data_ptr(),
                                  self.g_lp4.data_ptr(), _s()), 'rnr_tail_bwd')
        opt = self._opt if fused_opt else None
        comm = self.comm
        legacy_ar = self.allreduce if (ar_sum is None) else None
        has_early = self.early_layer in eng.wscratch_slices
        w0 = eng.wscratch_slices[self.early_layer][0] if has_early else 0
        overlap = eng.wscratch is not None and len(eng.wscratch_slices) == len(eng.specs) and has_early and \
            (fused_opt or ar_sum is not None or legacy_ar is not None)

        def after_layer(name):
            if not overlap or name != self.early_layer:
                return
            comm.wait_stream(main)                  # every weight gradient from `early_layer` to the last layer is complete
            with torch.cuda.stream(comm):
                if ar_sum is not None:
                    ar_sum(eng.wscratch[w0:])
                elif legacy_ar is not None:
                    legacy_ar(eng.wscratch[w0:])
                if opt is not None:
                    # Adam on 95 % of the parameters + their next-step GEMM matrices, underneath the rest of the backward pass
                    self._adam(opt['plan_early'])
                    eng.run_wprep_plan(opt['wprep_early'])

        def before_unpack():
            if not overlap:
                return
            if ar_sum is not None or legacy_ar is not None:
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    if w0 > 0:
                        (ar_sum or legacy_ar)(eng.wscratch[:w0])
                main.wait_stream(comm)

        gi = eng._backward_layers(after_layer, before_unpack, unpack=not fused_opt, input_grad_add=self.g_alb)
        if opt is not None:
            if not overlap:                        # (no early group on a second stream: everything here)
                self._adam(opt['plan_early'])
                eng.run_wprep_plan(opt['wprep_early'])
            self._adam(opt['plan_late'])
            eng.run_wprep_plan(opt['wprep_late'])
        gp = self._gtex_ptrs
        _lib.check(L.rnr_texmap_bwd(C.cast(gp, _pp), C.cast(self._tex_sizes, _ip), len(self.tex_grads), self.C,
                                    _cf(view['uv_map']).data_ptr(), _cf(view['sh_basis_map']).data_ptr(), 6, gi.data_ptr(), N, H, W, _s()),
                   'rnr_texmap_bwd')
        # envmap gradient -> SH coefficients: grad_coeff += basis_recon^T g_lp   (a17-bwd); the accumulator is cleared on the way
        lm = p.lighting_model
        _lib.check(L.rnr_sh_project_ld(lm.basis_val_recon.data_ptr(), self.g_lp4.data_ptr(), self._coeff_row_ptr(self.coeff_grad),
                                       self.Hl * self.Wl, self.B, 3, 4, 1.0, 1, _s()), 'rnr_sh_project_ld')
        main.wait_stream(side)                      # small-loss values and gradients are in place
        # loss value (device scalar; no host sync)
        cnt = float(N * 3 * (H - 2 * self.CROP) * (W - 2 * self.CROP))
        _lib.check(L.rnr_loss_combine(self.sums.data_ptr(), cnt, float(self.R), float(p.w['rays_lt_chrom']), self.sums.data_ptr() + 8 * 4, 2,
                                      self.loss_out.data_ptr(), int(self.sums.numel()), _s()), 'rnr_loss_combine')
        self._sums_clean = True                     # (loss_combine zeroed the accumulators behind itself)
        loss = self.loss_out[0]
        if ar_sum is not None or legacy_ar is not None:
            ar = ar_sum or legacy_ar
            if overlap:
                # weights are reduced already (GEMM-order scratch); what is left: biases / BatchNorm affine, then textures + SH
                # coefficients as ONE contiguous buffer
                ar(eng.grad_flat[eng.grad_small_offset:])
                ar(self.aux_flat)
            else:
                for g in self.grad_tensors():
                    ar(g)
            if ar_sum is not None and not fused_opt:
                for g in self.grad_tensors():
                    g.mul_(1.0 / float(self.world))
        if self.grad_hook is not None:
            self.grad_hook(self.grad_tensors())
        if opt is not None:
            self._adam(opt['plan_small'], advance=True)
            main.wait_stream(comm)
            self._mark_weights_current()            # the GEMM matrices were re-derived from the updated parameters
            self._grads_clean = True                # every accumulator was re-zeroed by the pass that consumed it
        else:
            self._grads_clean = False
            if step_optimizer:
                p.optimizer.step()
        return loss, self.final

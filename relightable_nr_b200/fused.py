"""The RNR training / rendering step with its per-pixel stages fused around the U-Net (csrc/fused.cu).

Same computation as ``RNRPipeline.forward / losses / backward`` (train_rnr.py:512-623) -- those drive the drop-in modules
one operator at a time, exactly like the reference script, and remain the parity reference for this file -- but:

* ``rnr_head_fwd`` writes the first convolution's operand (fp16 channels-last, reflect halo) directly from the texture
  pyramid and the per-view maps: no [N,108,H,W] fp32 tensor, no permute / cat / pack;
* ``rnr_tail_fwd / rnr_tail_bwd`` read the last convolution's NHWC output and write the data-gradient operand (bf16, zero
  halo), the bias gradient, the albedo and envmap gradients: no [N,26,3,H,W] temporaries;
* weight preparation for the data-gradient kernels, the two small losses (lighting L1, albedo mean; a few hundred tiny
  launches) and the gradient zero-fills run on a side stream underneath the U-Net forward;
* parameter gradients live in the engine's flat buffer; ``param.grad`` are views of it (no copies, one all-reduce bucket).

No autograd graph is built for the main path: the backward is the explicit kernel sequence below.  The small losses
still use torch autograd on ``textures`` / ``coeff`` (their cost is launch latency, hidden on the side stream).
"""
import ctypes as C

import torch

from . import _lib, ops
from .dropin import sph_harm as _sph_harm

vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
_pp = C.POINTER(C.c_void_p)
_ip = C.POINTER(C.c_int)
_lib.register_sigs({
    "rnr_head_fwd": [_pp, _ip, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, i32, i32, vp],
    "rnr_tail_fwd": [vp, i32, vp, vp, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "rnr_tail_bwd": [vp, i32, vp, vp, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, f32, f32, vp, i32, vp, vp, vp, vp],
})


def _s():
    return torch.cuda.current_stream().cuda_stream


def _cf(t):
    """fp32 contiguous view of a per-view map (a no-op for the maps ViewDataset / synthetic_view produce)."""
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class FusedRNRStep:
    """Fused execution of one RNRPipeline iteration.  Built lazily per (N, H, W); owns only scratch buffers."""

    CROP = 5

    def __init__(self, pipe):
        self.pipe = pipe
        self.dev = pipe.device
        self.L = _lib.lib()
        self._shape = None
        self.side = torch.cuda.Stream(device=self.dev)
        self.grad_hook = None          # callable(list of gradient tensors) between backward and the optimiser (data parallel)
        #: data-parallel averaging, overlapped with the backward pass: callable(tensor) that all-reduces ONE gradient buffer in
        #: place and scales it by 1/world (enqueued on the current stream).  The weight gradients of the layers that finish
        #: first in backward order -- 95 % of all parameters -- go out, still in GEMM order, while the full-resolution layers
        #: are being differentiated; the rest follows before the un-transpose, the small tensors and textures at the end.
        self.allreduce = None
        self.comm = torch.cuda.Stream(device=self.dev)
        self.early_layer = 'b3.down1'  # last layer (in backward order) whose weight gradient joins the early bucket

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
        self.sums = torch.zeros(4, dtype=torch.float64, device=dev)
        lm = p.lighting_model
        self.Hl, self.Wl = int(lm.lp_recon_h), int(lm.lp_recon_w)
        if need_backward:
            self.g_alb = torch.empty((N, 6, H, W), **f32k)
            self.g_lp4 = torch.zeros((self.Hl * self.Wl, 4), **f32k)
            # parameter gradients: persistent views (U-Net: into the engine's flat buffer)
            eng = self.eng
            seen = set()
            for k, prm in runner._unet().named_parameters(remove_duplicate=False):
                if k in eng.grad_slices and id(prm) not in seen:
                    prm.grad = eng.grad_view(k)
                    seen.add(id(prm))
            self.tex_grads = [torch.zeros_like(t) for t in tm.textures]
            for t, g in zip(tm.textures, self.tex_grads):
                t.grad = g
            self.coeff_grad = torch.zeros_like(lm.coeff)
            lm.coeff.grad = self.coeff_grad
        self._tex_ptrs = (C.c_void_p * len(tm.textures))(*[t.data_ptr() for t in tm.textures])
        self._tex_sizes = (C.c_int * len(tm.textures))(*[int(t.shape[1]) for t in tm.textures])
        self._tex_ids = [t.data_ptr() for t in tm.textures]
        self._shape = key

    def grad_tensors(self):
        """Every gradient buffer of the step (what a data-parallel all-reduce must average)."""
        return [self.eng.grad_flat] + list(self.tex_grads) + [self.coeff_grad]

    # ------------------------------------------------------------------------------------------------------------------
    def _head(self, view):
        p, eng = self.pipe, self.eng
        N, H, W = eng.N, eng.H, eng.W
        tm = p.texture_mapper
        if [t.data_ptr() for t in tm.textures] != self._tex_ids:
            raise RuntimeError('texture parameters were re-allocated; rebuild the fused step')
        cf = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
        _lib.check(self.L.rnr_head_fwd(
            C.cast(self._tex_ptrs, _pp), C.cast(self._tex_sizes, _ip), len(tm.textures), self.C,
            cf(view['uv_map']).data_ptr(), cf(view['sh_basis_map']).data_ptr(), 6,
            cf(view['TBN_map']).data_ptr(), cf(view['view_dir_map_tangent']).data_ptr(), cf(view['alpha_map']).data_ptr(),
            cf(view['normal_map']).data_ptr(), cf(view['view_dir_map']).data_ptr(),
            p.ray_sampler.pivots_dir.data_ptr(), self.Rs, p.ray_sampler_diffuse.pivots_dir.data_ptr(), self.Rd,
            eng.acts['input'].ptr, eng.acts_w['input'].ptr if eng.dual else None, eng.in_cpad,
            self.rays_uv.data_ptr(), self.albedo.data_ptr(), N, H, W, _s()), 'rnr_head_fwd')

    def _tail_fwd(self, view, lp):
        eng = self.eng
        raw = eng.layers['out'].raw
        _lib.check(self.L.rnr_tail_fwd(raw.data_ptr(), eng.out_ld, self.rays_uv.data_ptr(), self.albedo.data_ptr(), lp.data_ptr(),
                                       self.Hl, self.Wl, _cf(view['alpha_map']).data_ptr(), _cf(view['img_gt']).data_ptr(), self.Rs, self.Rd,
                                       eng.N, eng.H, eng.W, self.CROP, self.final.data_ptr(), self.aux.data_ptr(),
                                       self.sums.data_ptr(), _s()), 'rnr_tail_fwd')

    def _envmap(self):
        """LightingSH.reconstruct_lp (network.py:622-627) of the current coefficients as [Hl*Wl, 4] texels (r, g, b, 0): the SH
        coefficients get a zero fourth column, so the tail kernels fetch one 16-byte texel per bilinear tap."""
        lm = self.pipe.lighting_model
        with torch.no_grad():
            coeff4 = torch.nn.functional.pad(lm.coeff[self.pipe.lighting_idx], (0, 1))
            return _sph_harm.reconstruct_sh(coeff4, lm.basis_val_recon).contiguous()

    # ------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def render(self, view):
        """test_rnr.py:335-371 for one view: returns the final image [N,3,H,W] (valid until the next call)."""
        N, H, W = view['alpha_map'].shape
        self._setup(N, H, W, False)
        eng = self.eng
        training_bn, drop = self.runner.step_state(eng)
        lp = self._envmap()
        self._head(view)
        eng.forward(training=training_bn, drop_masks=drop)
        self.sums.zero_()
        self._tail_fwd(view, lp)
        return self.final

    def _small_losses(self):
        """lighting L1 (train_rnr.py:558-579) + albedo-mean loss (train_rnr.py:596-608) through torch autograd; their
        gradients accumulate into the (pre-zeroed) texture / coefficient gradient buffers."""
        p = self.pipe
        with torch.enable_grad():
            coeff = p.lighting_model.get_lighting_params(p.lighting_idx)
            l_est = _sph_harm.reconstruct_sh(coeff, p.lighting_model.basis_val)
            m = p.l_samples_init_mask.float()[:, None]
            d = (p.l_samples_init - l_est).abs()
            loss_lighting = (d * m).sum() / m.sum() * p.w['lighting'] + (d * (1 - m)).sum() / (1 - m).sum() * p.w['lighting_uncovered']
            tm = p.texture_mapper
            loss_alb = 0
            for c0 in (3, 0):
                tex = tm.flatten_mipmap(start_ch=c0, end_ch=c0 + 3)
                valid = (tex != tm.tex_flatten_mipmap_init[..., c0:c0 + 3]).any(dim=-1, keepdim=True).to(tex.dtype)
                cnt = valid.sum(dim=(0, 1, 2))
                loss_alb = loss_alb + ((tex * valid).sum(dim=(0, 1, 2)) / cnt.clamp(min=1) - 0.5).abs().sum() / 3 * (cnt > 0).float()
            small = loss_lighting + loss_alb * p.w['alb']
            small.backward()
        # autograd accumulates in place into the pre-set .grad buffers; a replaced tensor would silently detach them
        for t, g in zip(tm.textures, self.tex_grads):
            assert t.grad is g, 'texture gradient buffer was replaced'
        assert p.lighting_model.coeff.grad is self.coeff_grad, 'coefficient gradient buffer was replaced'
        return small.detach()

    def train_step(self, view, step_optimizer=True):
        """One iteration (train_rnr.py:490-623): returns (loss, final image)."""
        p = self.pipe
        N, H, W = view['alpha_map'].shape
        self._setup(N, H, W, True)
        eng, L = self.eng, self.L
        main = torch.cuda.current_stream(self.dev)
        side = self.side
        training_bn, drop = self.runner.step_state(eng)

        # ---- side stream: weight matrices, gradient zero-fills, the two small losses ----
        side.wait_stream(main)
        with torch.cuda.stream(side):
            eng.prepare_weights_split('fwd')
            ev_w = torch.cuda.Event()
            ev_w.record(side)
            eng.zero_grads()
            self.g_lp4.zero_()
            for g in self.tex_grads:
                g.zero_()
            self.coeff_grad.zero_()
            ev_z = torch.cuda.Event()
            ev_z.record(side)
            eng.prepare_weights_split('dgrad')
            small = self._small_losses()
        # ---- main stream: envmap, head, U-Net forward, tail ----
        self.sums.zero_()
        lp = self._envmap()
        self._head(view)
        main.wait_event(ev_w)
        eng.forward(training=training_bn, drop_masks=drop, weights_ready=True)
        self._tail_fwd(view, lp)
        main.wait_event(ev_z)
        sp = eng.specs[-1]
        raw = eng.layers['out'].raw
        _lib.check(L.rnr_tail_bwd(raw.data_ptr(), eng.out_ld, self.rays_uv.data_ptr(), self.albedo.data_ptr(), lp.data_ptr(),
                                  self.Hl, self.Wl, _cf(view['alpha_map']).data_ptr(), _cf(view['img_gt']).data_ptr(), self.Rs, self.Rd,
                                  N, H, W, self.CROP, self.aux.data_ptr(), self.sums.data_ptr(), 1.0, float(p.w['rays_lt_chrom']),
                                  eng.gz['out'].ptr, eng.out_ld, eng.grad_view(sp.b_key).data_ptr(), self.g_alb.data_ptr(),
                                  self.g_lp4.data_ptr(), _s()), 'rnr_tail_bwd')
        main.wait_stream(side)                      # data-gradient weight matrices + small-loss gradients are in place
        after_layer = before_unpack = None
        overlap = (self.allreduce is not None and eng.wscratch is not None and len(eng.wscratch_slices) == len(eng.specs)
                   and self.early_layer in eng.wscratch_slices)
        if overlap:
            comm, ar = self.comm, self.allreduce
            w0 = eng.wscratch_slices[self.early_layer][0]

            def after_layer(name):
                if name != self.early_layer:
                    return
                comm.wait_stream(main)              # every weight gradient from `early_layer` to the last layer is complete
                with torch.cuda.stream(comm):
                    ar(eng.wscratch[w0:])

            def before_unpack():
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    if w0 > 0:
                        ar(eng.wscratch[:w0])
                main.wait_stream(comm)
        gi = eng._backward_layers(after_layer, before_unpack)   # [N, C, H, W] gradient w.r.t. the texture channels of the input
        gi[:, :6] += self.g_alb
        tm = p.texture_mapper
        gp = (C.c_void_p * len(self.tex_grads))(*[g.data_ptr() for g in self.tex_grads])
        _lib.check(L.rnr_texmap_bwd(C.cast(gp, _pp), C.cast(self._tex_sizes, _ip), len(self.tex_grads), self.C,
                                    _cf(view['uv_map']).data_ptr(), _cf(view['sh_basis_map']).data_ptr(), 6, gi.data_ptr(), N, H, W, _s()),
                   'rnr_texmap_bwd')
        # envmap gradient -> SH coefficients: grad_coeff += basis_recon^T g_lp   (a17-bwd)
        lm = p.lighting_model
        g_lp = self.g_lp4[:, :3].contiguous()
        _lib.check(L.rnr_sh_project(lm.basis_val_recon.data_ptr(), g_lp.data_ptr(),
                                    self.coeff_grad[p.lighting_idx].data_ptr(), g_lp.shape[0], lm.basis_val_recon.shape[1], 3, 1, 1.0,
                                    _s()), 'rnr_sh_project')
        # loss value (device scalars; no host sync)
        cnt = float(N * 3 * (H - 2 * self.CROP) * (W - 2 * self.CROP))
        loss = (self.sums[2] / cnt + self.sums[0] / self.sums[1] / self.R * p.w['rays_lt_chrom']).float() + small
        if overlap:
            # weights are averaged already (GEMM-order scratch); what is left: biases / BatchNorm affine, textures, SH coefficients
            self.allreduce(eng.grad_flat[eng.grad_small_offset:])
            for g in self.tex_grads:
                self.allreduce(g)
            self.allreduce(self.coeff_grad)
        elif self.allreduce is not None:
            for g in self.grad_tensors():
                self.allreduce(g)
        if self.grad_hook is not None:
            self.grad_hook(self.grad_tensors())
        if step_optimizer:
            p.optimizer.step()
        return loss, self.final

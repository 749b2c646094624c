"""The fused RNR step (relightable_nr_b200/fused.py: head / tail kernels around the U-Net, explicit backward) against
(a) the same step driven operator by operator through the drop-in modules and (b) the CPU oracle of train_rnr.py:512-608.

Tolerances: the head kernel's operand equals the module path's packed operand bit for bit (same arithmetic, same fp16
rounding); final image max-abs <= 1e-5 and PSNR >= 50 dB vs the oracle; gradients cosine >= 0.98 vs the fp32 oracle and
>= 0.999 vs the module path (atomic accumulation order is the only difference)."""
import pytest
import torch

from tests.util import cosine, psnr

pytestmark = pytest.mark.gpu


def _pipe(**kw):
    from relightable_nr_b200.pipeline import RNRPipeline
    cfg = dict(device='cuda:0', img_size=64, texture_size=64, texture_num_ch=24, mipmap_level=3, nf0=16, sh_lmax=4,
               num_l_samples=512, lp_recon_h=16, lp_recon_w=32, dropout=False)
    cfg.update(kw)
    return RNRPipeline(**cfg)


def test_head_operand_is_bit_identical_to_the_module_path():
    from relightable_nr_b200.pipeline import synthetic_view
    pipe = _pipe()
    view = synthetic_view(64, view_idx=5, device='cuda:0')
    with torch.no_grad():
        pipe.forward(view)                                  # module path: texture mapper, ray samplers, cat, pack
    f = pipe.fused
    f._setup(1, 64, 64, False)
    eng_m = [e for e in pipe.render_net.net._runner._engines.values()][0]
    ref = eng_m.acts['input'].t.clone()
    with torch.no_grad():
        f.render(view)
    got = f.eng.acts['input'].t
    assert got.shape == ref.shape
    assert torch.equal(got, ref), (got.float() - ref.float()).abs().max().item()
    # rays_uv of the two samplers, concatenated
    a = view['alpha_map'][:, None].permute(0, 2, 3, 1)
    uv0 = pipe.ray_sampler(view['TBN_map'], view['view_dir_map_tangent'], a)[1]
    uv1 = pipe.ray_sampler_diffuse(view['TBN_map'], view['view_dir_map_tangent'], a)[1]
    assert torch.equal(f.rays_uv, torch.cat((uv0, uv1), -1))


def test_fused_render_matches_module_path_and_oracle():
    from oracle.rnr_step import rnr_step, state_from_pipeline
    from relightable_nr_b200.pipeline import synthetic_view
    pipe = _pipe()
    view = synthetic_view(64, view_idx=5, device='cuda:0')
    state = state_from_pipeline(pipe)
    mod = pipe.render(view).clone()
    fus = pipe.render(view, fused=True).clone()
    assert (mod - fus).abs().max().item() <= 1e-5
    _, ref_final, _ = rnr_step(state, view, requires_grad=False)
    assert psnr(fus.cpu(), ref_final) >= 50.0


def test_fused_step_gradients_match_module_path_and_oracle():
    from oracle.rnr_step import rnr_step, state_from_pipeline
    from relightable_nr_b200.pipeline import synthetic_view
    view = synthetic_view(64, view_idx=5, device='cuda:0')
    pipe = _pipe()
    state = state_from_pipeline(pipe)
    # module path gradients
    final, rays_lt, alpha = pipe.forward(view)
    loss_m, _ = pipe.losses(view, final, rays_lt, alpha)
    loss_m.backward()
    gm = {'tex%d' % i: t.grad.clone() for i, t in enumerate(pipe.texture_mapper.textures)}
    gm['coeff'] = pipe.lighting_model.coeff.grad.clone()
    gm_net = {k: p_.grad.clone() for k, p_ in pipe.render_net.named_parameters() if p_.grad is not None}
    pipe.optimizer.zero_grad(set_to_none=True)
    # fused path, same parameters
    loss_f, final_f = pipe.fused.train_step(view, step_optimizer=False)
    torch.cuda.synchronize()
    assert abs(loss_f.item() - loss_m.item()) <= 1e-4 * max(1.0, abs(loss_m.item())), (loss_f.item(), loss_m.item())
    assert (final_f - final.detach()).abs().max().item() <= 1e-5
    for i, t in enumerate(pipe.texture_mapper.textures):
        c = cosine(t.grad, gm['tex%d' % i])
        print('texture %d fused-vs-module cosine %.6f' % (i, c))
        assert c >= 0.999
    assert cosine(pipe.lighting_model.coeff.grad, gm['coeff']) >= 0.999
    worst = 1.0
    n = 0
    for k, p_ in pipe.render_net.named_parameters():
        if k in gm_net:
            assert p_.grad is not None, k
            worst = min(worst, cosine(p_.grad, gm_net[k]))
            n += 1
    print('worst U-Net grad cosine fused-vs-module %.6f over %d tensors' % (worst, n))
    assert n > 40 and worst >= 0.999
    assert not any(p_.grad is not None for k, p_ in pipe.render_net.named_parameters() if '.fuse.' in k)
    # oracle
    ref_loss, ref_final, ref_grads = rnr_step(state, view)
    assert abs(loss_f.item() - ref_loss.item()) <= 2e-3 * max(1.0, abs(ref_loss.item()))
    assert psnr(final_f.cpu(), ref_final) >= 50.0
    for i, t in enumerate(pipe.texture_mapper.textures):
        assert cosine(t.grad.cpu(), ref_grads['textures.%d' % i]) >= 0.98
    assert cosine(pipe.lighting_model.coeff.grad[0].cpu(), ref_grads['coeff']) >= 0.98
    worst = 1.0
    for k, p_ in pipe.render_net.named_parameters():
        if p_.grad is not None and 'unet/' + k in ref_grads:
            worst = min(worst, cosine(p_.grad.cpu(), ref_grads['unet/' + k]))
    print('worst U-Net grad cosine fused-vs-oracle %.5f' % worst)
    assert worst >= 0.98


def test_fused_training_trajectory_matches_module_path():
    """6 optimiser steps, dropout off: same loss trajectory (2e-3 relative), untouched texels stay bit-identical."""
    from relightable_nr_b200.pipeline import synthetic_view
    views = [synthetic_view(64, view_idx=i, device='cuda:0') for i in (2, 9, 4)]
    a, b = _pipe(), _pipe()
    before = b.texture_mapper.textures[0].detach().clone()
    ref = [a.train_step(v)[0].item() for v in views for _ in range(2)]
    got = [b.train_step(v, fused=True)[0].item() for v in views for _ in range(2)]
    print(ref, got)
    for x, y in zip(ref, got):
        assert abs(x - y) <= 2e-3 * max(1.0, abs(x)), (ref, got)
    t0 = b.texture_mapper.textures[0].detach()
    changed = (t0[..., 6:] != before[..., 6:]).any(-1)[0]
    assert changed.any() and not changed.all()


def test_fused_cuda_graph_step():
    from relightable_nr_b200.pipeline import synthetic_view
    views = [synthetic_view(64, view_idx=i, device='cuda:0') for i in (2, 9, 4)]
    eager = _pipe(capturable=True)
    eager.train_step(views[0], fused=True)
    ref = [eager.train_step(v, fused=True)[0].item() for v in views for _ in range(2)]
    pipe = _pipe(capturable=True)
    step, static = pipe.make_graphed_step(views[0], warmup=1, fused=True)
    assert pipe.graph_launches > 80
    got = [float(step(v)) for v in views for _ in range(2)]
    print(ref, got)
    for x, y in zip(ref, got):
        assert abs(x - y) <= 2e-3 * max(1.0, abs(x)), (ref, got)


@pytest.mark.parametrize('size,nf0', [(64, 16), (128, 16), (64, 64)])
def test_fused_training_trajectory_follows_the_oracle(size, nf0):
    """(nf0 = 64 runs every layer on the halo kernel: fused BatchNorm finalize, fused un-transpose + Adam, early optimiser group.)
    20 consecutive Adam iterations (3 views cycled, dropout off): the fused GPU step (fp16 activations, bf16 gradients) against
    the CPU fp32 oracle running the SAME iterations with torch.optim.Adam (oracle.rnr_step.rnr_trajectory).  Gate: the loss
    agrees within 1 % at EVERY step, and the accumulated parameter updates of the textures / SH coefficients point the same
    way (cosine >= 0.9) -- the evidence that reduced-precision gradients train the same model."""
    from oracle.rnr_step import rnr_trajectory, state_from_pipeline
    from relightable_nr_b200.pipeline import synthetic_view
    pipe = _pipe(img_size=size, nf0=nf0)
    views = [synthetic_view(size, view_idx=i, device='cuda:0') for i in (2, 9, 4)]
    seq = [views[i % 3] for i in range(20)]
    state = state_from_pipeline(pipe)
    got = [pipe.train_step(v, fused=True)[0].item() for v in seq]
    torch.cuda.synchronize()
    ref, end = rnr_trajectory(state, seq, lr=1e-3)
    dev = [abs(a - b) / max(abs(b), 1e-12) for a, b in zip(got, ref)]
    print('loss GPU   :', ' '.join('%.5f' % x for x in got))
    print('loss oracle:', ' '.join('%.5f' % x for x in ref))
    print('max relative deviation over 20 steps: %.3e' % max(dev))
    assert ref[-1] < ref[0], 'the oracle trajectory must make progress for the comparison to mean anything'
    assert max(dev) <= 1e-2, dev
    for i, t in enumerate(pipe.texture_mapper.textures):
        d_gpu = t.detach().cpu() - state['textures'][i]
        d_ref = end['textures'][i] - state['textures'][i]
        c = cosine(d_gpu, d_ref)
        print('texture %d accumulated-update cosine %.4f' % (i, c))
        assert c >= 0.9
    c = cosine(pipe.lighting_model.coeff.detach()[0].cpu() - state['coeff'], end['coeff'] - state['coeff'])
    print('SH coefficient accumulated-update cosine %.4f' % c)
    assert c >= 0.9


@pytest.mark.parametrize('size,nf0,early', [(64, 64, 'down3'), (64, 16, 'down3'), (128, 64, None)])
def test_optimiser_pass_writes_the_next_steps_gemm_matrices_bit_exactly(size, nf0, early):
    """The fused un-transpose + Adam kernel also emits next step's 16-bit forward / data-gradient GEMM matrices (csrc/optim.cu
    phases 3-4).  After three optimiser steps every matrix must equal, bit for bit, what the weight-preparation kernel derives
    from the updated fp32 parameters -- for the families the optimiser pass covers AND the ones left to the preparation plan."""
    from relightable_nr_b200.pipeline import synthetic_view
    pipe = _pipe(img_size=size, nf0=nf0)
    views = [synthetic_view(size, view_idx=i, device='cuda:0') for i in (2, 9, 4)]
    for v in views:
        pipe.train_step(v, fused=True)
    torch.cuda.synchronize()
    f = pipe.fused
    eng = f.eng
    in_adam = f._opt['gemm_in_adam']
    print('forward matrices written by the optimiser pass:', in_adam[0])
    print('data-gradient matrices written by the optimiser pass:', in_adam[1])
    if nf0 == 64:
        assert len(in_adam[0]) >= len(eng.specs) - 2, 'the 64-channel-chunk layers are expected to be covered'
    mats = []
    for sp in eng.specs:
        st = eng.layers[sp.name]
        for kind, jobs in (('fwd', st.wprep_fwd), ('dgrad', st.wprep_dgrad)):
            for j, w in enumerate(jobs):
                mats.append(('%s/%s/%d' % (sp.name, kind, j), w.dst))
        for kind, d in (('fwd', st.wmat_fwd), ('dgrad', st.wmat_dgrad)):
            if d is not None:
                mats.append(('%s/%s/all' % (sp.name, kind), d['base']))
    got = [(n, t.clone()) for n, t in mats]
    eng.prepare_weights(backward=True)
    torch.cuda.synchronize()
    bad = [n for (n, a), (_, t) in zip(got, mats) if not torch.equal(a.view(torch.int16), t.view(torch.int16))]
    assert not bad, bad


def test_full_size_fused_step_matches_oracle():
    """BASELINE.json's benchmark configuration itself (512x512 view, texture 512^2 x 24 ch x 4 mips, U-Net 108 -> 78 with nf0 = 64,
    26 rays, SH lmax 10 / 256x512 envmap): one fused training step on the GPU against the CPU fp32 oracle of train_rnr.py:512-608.
    Gates: render PSNR >= 50 dB, loss within 2e-3 relative, gradient cosine >= 0.98 for the textures / SH coefficients and
    >= 0.97 for every U-Net tensor (fp16 activations, bf16 gradients; ~0.1 % of the ReLU gates flip at this depth)."""
    from oracle.rnr_step import rnr_step, state_from_pipeline
    from relightable_nr_b200.pipeline import RNRPipeline, synthetic_view
    pipe = RNRPipeline(device='cuda:0', img_size=512, dropout=False)
    view = synthetic_view(512, view_idx=7, device='cuda:0')
    state = state_from_pipeline(pipe)
    loss, final = pipe.fused.train_step(view, step_optimizer=False)
    torch.cuda.synchronize()
    ref_loss, ref_final, ref_grads = rnr_step(state, view)
    p = psnr(final.cpu(), ref_final)
    print('512^2 fused step: PSNR %.1f dB, loss %.6f vs oracle %.6f' % (p, loss.item(), ref_loss.item()))
    assert p >= 50.0
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * max(1.0, abs(ref_loss.item()))
    for i, t in enumerate(pipe.texture_mapper.textures):
        c = cosine(t.grad.cpu(), ref_grads['textures.%d' % i])
        print('texture %d grad cosine %.5f' % (i, c))
        assert c >= 0.98
    assert cosine(pipe.lighting_model.coeff.grad[0].cpu(), ref_grads['coeff']) >= 0.98
    worst, wk = 1.0, None
    for k, p_ in pipe.render_net.named_parameters():
        if p_.grad is not None and 'unet/' + k in ref_grads:
            c = cosine(p_.grad.cpu(), ref_grads['unet/' + k])
            if c < worst:
                worst, wk = c, k
    print('worst U-Net grad cosine %.5f (%s)' % (worst, wk))
    assert worst >= 0.97
